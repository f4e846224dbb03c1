// Batched recursive_unescape / recursive_escape on strings for sm_100a.
//
// Replaces the string arms of interp.rs:147-161 / :163-177 and the inline replaces of `print` and
// `write` (runtime.rs:1053-1055, 1272).  The reference runs two sequential str::replace passes;
// both compose to one streaming pass:
//   unescape: drop a '\' exactly when the next byte is '{' or '}'.  The first pass ("\{" -> "{")
//             cannot create a new "\}" (a surviving '\' is then followed by '{'), so the second pass
//             sees exactly the original "\}" pairs.
//   escape:   prefix every '{' and every '}' with '\' (the two passes touch disjoint bytes).
// One thread per string, IE_TILE strings per CTA, compacted output via the shared tile scan.
#include <cuda_runtime.h>

#include "ie_kernels.h"
#include "ie_scan.cuh"

namespace {

template <bool WRITE>
__device__ __forceinline__ uint64_t transform(int mode, const uint8_t* __restrict__ s, uint64_t n, uint8_t* __restrict__ dst) {
    uint64_t o = 0;
    if (mode == 0) {
        for (uint64_t i = 0; i < n; ++i) {
            const uint8_t c = __ldg(s + i);
            if (c == '\\' && i + 1 < n) {
                const uint8_t d = __ldg(s + i + 1);
                if (d == '{' || d == '}') continue;  // the brace itself is copied on the next iteration
            }
            if (WRITE) dst[o] = c;
            ++o;
        }
    } else {
        for (uint64_t i = 0; i < n; ++i) {
            const uint8_t c = __ldg(s + i);
            if (c == '{' || c == '}') { if (WRITE) dst[o] = '\\'; ++o; }
            if (WRITE) dst[o] = c;
            ++o;
        }
    }
    return o;
}

__global__ void __launch_bounds__(IE_TILE) ie_escape_kernel(int mode, const uint8_t* __restrict__ in, const uint64_t* __restrict__ offs,
                                                            uint64_t n, uint8_t* __restrict__ out, uint64_t out_cap,
                                                            uint64_t* __restrict__ out_offs, IeWorkspace ws) {
    __shared__ ie_scan::TileSmem s_scan;
    const uint32_t tile = ie_scan::acquire_tile(s_scan, ws.tile_counter);
    const uint64_t i = (uint64_t)tile * IE_TILE + threadIdx.x;
    const bool active = i < n;
    const uint8_t* s = nullptr;
    uint64_t len = 0, olen = 0;
    if (active) {
        const uint64_t a = __ldg(offs + i);
        len = __ldg(offs + i + 1) - a;
        s = in + a;
        olen = transform<false>(mode, s, len, nullptr);
    }
    uint64_t tile_end;
    const uint64_t off = ie_scan::exclusive_prefix(s_scan, ws.tile_state, tile, olen, &tile_end);
    if (threadIdx.x == 0 && (uint64_t)tile + 1 == (n + IE_TILE - 1) / IE_TILE) out_offs[n] = tile_end;
    if (!active) return;
    out_offs[i] = off;
    if (off + olen > out_cap) { *ws.overflow = 1u; return; }
    transform<true>(mode, s, len, out + off);
}

}  // namespace

cudaError_t ie_launch_escape(int mode, const uint8_t* d_in, const uint64_t* d_in_offs, uint64_t n, uint8_t* d_out,
                             uint64_t out_cap, uint64_t* d_out_offs, const IeWorkspace& ws, cudaStream_t stream) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(ws.zero_base, 0, ws.zero_bytes, stream)) != cudaSuccess) return err;
    if (n == 0) return cudaMemsetAsync(d_out_offs, 0, sizeof(uint64_t), stream);
    const uint64_t tiles = (n + IE_TILE - 1) / IE_TILE;
    ie_escape_kernel<<<(unsigned)tiles, IE_TILE, 0, stream>>>(mode, d_in, d_in_offs, n, d_out, out_cap, d_out_offs, ws);
    return cudaGetLastError();
}
