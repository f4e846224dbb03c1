// Batched recursive_unescape / recursive_escape on strings for sm_100a.
//
// Replaces the string arms of interp.rs:147-161 / :163-177 and the inline replaces of `print` and
// `write` (runtime.rs:1053-1055, 1272).  The reference runs two sequential str::replace passes;
// both compose to one streaming pass:
//   unescape: drop a '\' exactly when the next byte OF THE SAME STRING is '{' or '}'.  The first pass
//             ("\{" -> "{") cannot create a new "\}" (a surviving '\' is then followed by '{'), so the
//             second pass sees exactly the original "\}" pairs.
//   escape:   prefix every '{' and every '}' with '\' (the two passes touch disjoint bytes).
//
// The arena is processed as ONE flat byte stream, not string by string: a CTA tile is 8 KiB of
// consecutive arena bytes (512 aligned 16-byte chunks, two per thread, coalesced).  Every chunk gets a
// 16-bit mask of its special bytes by SIMD-in-register compares; a CTA scan of the per-chunk output
// sizes plus a decoupled look-back over tiles gives every byte its output position in a single pass over
// HBM.  The kept / inserted bytes are pushed into a bank-conflict-free staging buffer in shared memory
// (one pad word per 16 bytes) laid out congruent to the output address modulo 16, and leave the SM as
// 16-byte coalesced stores.  out_offs[i] is the image of in_offs[i] under the same position map; the
// strings that start inside a tile come from a small pre-pass over the offset array (tile_first[]).
#include <cuda_runtime.h>

#include "ie_kernels.h"
#include "ie_scan.cuh"

#ifdef IE_PHASE_TIMING
__device__ unsigned long long g_esc_cycles[16];
#define PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_esc_cycles[k], (unsigned long long)(now_ - t_phase_)); t_phase_ = now_; } } while (0)
#define PHASE_INIT() long long t_phase_ = clock64()
__device__ unsigned long long g_esc_trace[4 * 32768];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TRACE(tile, k) do { if (threadIdx.x == 0 && (tile) < 32768) g_esc_trace[4 * (tile) + (k)] = gtime(); } while (0)
extern "C" int ie_debug_escape_trace(unsigned long long* out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_esc_trace, sizeof(g_esc_trace)) != cudaSuccess;
}
extern "C" int ie_debug_escape_cycles(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, g_esc_cycles, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    cudaMemcpyFromSymbol(out16 + 12, ie_scan::g_lb_stats, sizeof(unsigned long long) * 4);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_esc_cycles, z, sizeof z); cudaMemcpyToSymbol(ie_scan::g_lb_stats, z, 32); }
    return 0;
}
#else
#define PHASE_MARK(k) do { } while (0)
#define PHASE_INIT() do { } while (0)
#define TRACE(tile, k) do { } while (0)
#endif

namespace {

constexpr int NT = 256;
#ifndef IE_ESC_CTAS
#define IE_ESC_CTAS 8
#endif
constexpr int ESC_CTAS = IE_ESC_CTAS;  // persistent CTAs per SM
constexpr int CPT = 2;                      // chunks per thread
constexpr int TILE_CHUNKS = NT * CPT;       // 512
constexpr int TILE_BYTES = TILE_CHUNKS * 16;
static_assert(TILE_BYTES == IE_ESCAPE_TILE_BYTES, "ie_escape_tiles() sizes the look-back state");
constexpr int STAGE_CHUNKS = TILE_CHUNKS * 5 / 4 + 2;  // staged output of a tile; brace-dense tiles that escape to more
                                                     // than 1.25x write their bytes straight to global memory

struct Smem {
    ie_scan::TileSmemT<NT> scan;
    uint32_t cpre[TILE_CHUNKS];     // output bytes of the tile before each chunk
    uint32_t cmask[TILE_CHUNKS];    // low 16: bytes of the chunk that belong to the arena; high 16: special bytes
    uint32_t fix[TILE_BYTES / 32];  // unescape: bit p = the byte after p starts a new string (a '\' at p stays)
    uint32_t warp_tot[NT / 32];
    uint64_t bound[2];              // strings [bound[0], bound[1]) start inside (or right after) this tile
    uint32_t stage[STAGE_CHUNKS * 5];
};

__device__ __forceinline__ uint32_t eqmask(uint32_t w, uint32_t pat) {  // 0x80 in every byte of w equal to pat's
    const uint32_t x = w ^ pat;
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t bits4(uint32_t m) {  // 0x80-per-byte flags -> 4 adjacent bits
    return ((m >> 7) * 0x00204081u >> 21) & 0xFu;
}
__device__ __forceinline__ uint32_t stage_addr(uint32_t b) { return b + ((b >> 4) << 2); }  // byte b of the padded layout

// First index i in [0, cnt] with i == cnt or (STRICT ? offs[i] > key : offs[i] >= key); whole warp.
template <bool STRICT>
__device__ __forceinline__ uint64_t warp_bound(const uint64_t* __restrict__ offs, uint64_t cnt, uint64_t key, uint32_t lane) {
    uint64_t lo = 0, hi = cnt;  // every i < lo fails the predicate, every i >= hi satisfies it
    while (hi - lo > 32) {
        const uint64_t step = (hi - lo) >> 5;
        const uint64_t v = __ldg(offs + lo + lane * step);
        const uint32_t below = __ballot_sync(0xFFFFFFFFu, STRICT ? v <= key : v < key);
        const uint32_t k = __popc(below);
        if (k == 0) { hi = lo; break; }
        const uint64_t base = lo;
        lo = base + (k - 1) * step + 1;
        if (k < 32) hi = base + k * step;
    }
    bool below = false;
    if (lo + lane < hi) { const uint64_t v = __ldg(offs + lo + lane); below = STRICT ? v <= key : v < key; }
    return lo + __popc(__ballot_sync(0xFFFFFFFFu, below));
}

// Pre-pass, one thread per string start (n + 1 of them), fully parallel:
//  * tile_first[t] = the first string that starts at or after tile t's first byte, so the streaming kernel
//    finds "its" strings with one load instead of a search of the offset array;
//  * unescape only: a '\' that is the LAST byte of its string stays even when the next string starts with a
//    brace.  The (rare) positions where that happens go to a list, so the streaming kernel can size a tile
//    without knowing the string starts.  More than IE_ESCAPE_FIX_CAP hits make it search per tile instead.
__global__ void __launch_bounds__(256) ie_escape_prepass_kernel(int mode, const uint8_t* __restrict__ in, const uint64_t* __restrict__ offs,
                                                                uint64_t n, uint64_t* __restrict__ tile_first, uint64_t max_tiles,
                                                                uint32_t* __restrict__ fix_count, uint64_t* __restrict__ fix_list) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint64_t base0 = __ldg(offs), end0 = __ldg(offs + n);
    const int64_t pa0 = (int64_t)base0 - (int64_t)((uintptr_t)(in + base0) & 15);
    const uint64_t s = __ldg(offs + i);
    // tiles whose first byte lies in (previous start, this start]
    const int64_t tb = ((int64_t)s - pa0) / TILE_BYTES;
    const int64_t ta = i == 0 ? -1 : ((int64_t)__ldg(offs + i - 1) - pa0) / TILE_BYTES;
    for (int64_t t = ta + 1; t <= tb && (uint64_t)t < max_tiles; ++t) tile_first[t] = i;
    if (mode != 0 || i == 0 || i >= n || s == base0 || s >= end0) return;
    const uint8_t b = __ldg(in + s);
    if ((b == '{' || b == '}') && __ldg(in + s - 1) == '\\') {
        const uint32_t k = atomicAdd(fix_count, 1u);
        if (k < IE_ESCAPE_FIX_CAP) fix_list[k] = s - 1;
    }
}

// A CTA owns a SPAN of `span_tiles` consecutive 8 KiB tiles.  Pass 1 streams the span once and only counts
// (all of its loads are independent: deep memory-level parallelism); the span total is published and the
// look-back runs ONCE per span; pass 2 re-reads the span (L2 hits), scans tile by tile and writes.  Only
// spans, not tiles, are ordered, so the time a CTA spends waiting for its predecessors is paid per 64 KiB.
__global__ void __launch_bounds__(NT, ESC_CTAS) ie_escape_kernel(int mode, const uint8_t* __restrict__ in, const uint64_t* __restrict__ offs,
                                                       uint64_t n, uint8_t* __restrict__ out, uint64_t out_cap,
                                                       uint64_t* __restrict__ out_offs, IeWorkspace ws, uint64_t max_tiles,
                                                       uint32_t span_tiles) {
    __shared__ Smem sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base0 = __ldg(offs), end0 = __ldg(offs + n);
    // the flat stream starts at the aligned floor of the first byte; positions are absolute arena indices
    const uint32_t lead = (uint32_t)((uintptr_t)(in + base0) & 15);
    const int64_t pa0 = (int64_t)base0 - lead;
    uint64_t tiles = (lead + (end0 - base0) + TILE_BYTES - 1) / TILE_BYTES;
    if (tiles == 0) tiles = 1;
    if (tiles > max_tiles) { if (tid == 0 && blockIdx.x == 0) *ws.overflow = 1u; return; }  // in_bytes understated
    const uint64_t spans = (tiles + span_tiles - 1) / span_tiles;
    const uint32_t n_fix = mode == 0 ? *ws.fix_count : 0u;  // written by the pre-pass earlier on this stream
    const bool use_list = n_fix <= IE_ESCAPE_FIX_CAP;
    uint8_t* stage = reinterpret_cast<uint8_t*>(sm.stage);

    uint4 v[CPT];
    uint32_t nb[CPT];  // the byte after the chunk
    uint32_t valid[CPT], special[CPT], cnt[CPT];

    // chunk c = 64 * warp + 32 * u + lane of the tile that starts at t0, coalesced
    // interior tile: every byte of it and the byte after it belong to the arena (no range checks per chunk)
    auto is_interior = [&](int64_t t0) { return t0 >= (int64_t)base0 && t0 + TILE_BYTES < (int64_t)end0; };
    auto load_tile = [&](int64_t t0) {
        const uint8_t* tp = in + t0 + (64 * warp + lane) * 16;
        if (is_interior(t0)) {
#pragma unroll
            for (int u = 0; u < CPT; ++u) {
                v[u] = __ldg(reinterpret_cast<const uint4*>(tp + u * 512));
                nb[u] = mode == 0 ? __ldg(tp + u * 512 + 16) : 0u;
            }
            return;
        }
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            const int64_t p = t0 + (int64_t)(64 * warp + 32 * u + lane) * 16;
            v[u] = make_uint4(0, 0, 0, 0);
            nb[u] = 0;
            if (p + 16 > (int64_t)base0 && p < (int64_t)end0) {
                v[u] = __ldg(reinterpret_cast<const uint4*>(in + p));
                if (mode == 0 && p + 16 < (int64_t)end0) nb[u] = __ldg(in + p + 16);
            }
        }
    };
    // strings starting in [t0, t0 + TILE_BYTES]: the closed end marks a '\' that ends a string at the tile's
    // last byte; the last tile also owns every start at end0 (trailing empty strings, offs[n])
    auto find_bounds = [&](int64_t t0, bool last) {
        if (warp == 0) {
            const uint64_t key = t0 > 0 ? (uint64_t)t0 : 0;
            const uint64_t b = warp_bound<false>(offs, n + 1, key, lane);
            if (lane == 0) sm.bound[0] = b;
        } else if (warp == 1) {
            const uint64_t b = last ? n + 1 : warp_bound<true>(offs, n + 1, (uint64_t)(t0 + TILE_BYTES), lane);
            if (lane == 0) sm.bound[1] = b;
        }
        __syncthreads();
    };
    // unescape: sm.fix bit p = the byte after p starts a new string.  Normally taken from the pre-pass list
    // (usually empty); with more than IE_ESCAPE_FIX_CAP hits every string start of the tile is marked instead.
    auto build_fix = [&](int64_t t0, bool last) -> bool {
        if (mode != 0 || (n_fix == 0 && use_list)) return false;
        __syncthreads();  // readers of the previous tile's bits are done
        sm.fix[tid] = 0;
        if (use_list) {
            __syncthreads();
            for (uint32_t j = tid; j < n_fix; j += NT) {
                const int64_t q = (int64_t)ws.fix_list[j] - t0;
                if (q >= 0 && q < TILE_BYTES) atomicOr(&sm.fix[q >> 5], 1u << (q & 31));
            }
        } else {
            find_bounds(t0, last);
            for (uint64_t i = sm.bound[0] + tid; i < sm.bound[1]; i += NT) {
                const int64_t q = (int64_t)__ldg(offs + i) - 1 - t0;  // tile-local position of the byte before the start
                if (q >= 0 && q < TILE_BYTES) atomicOr(&sm.fix[q >> 5], 1u << (q & 31));
            }
        }
        __syncthreads();
        return true;
    };
    // masks and output sizes of the loaded chunks
    auto classify = [&](int64_t t0, bool have_fix) {
        const bool interior = is_interior(t0);
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            const uint32_t c = 64 * warp + 32 * u + lane;
            int64_t b = 17;  // bytes of the arena from the chunk's first byte on (> 16: the next byte exists)
            valid[u] = 0xFFFFu;
            if (!interior) {
                const int64_t p = t0 + (int64_t)c * 16;
                const int64_t a = (int64_t)base0 - p;  // valid bytes are [a, b) of the chunk
                b = (int64_t)end0 - p;
                const uint32_t va = a <= 0 ? 0xFFFFu : (a >= 16 ? 0u : (0xFFFFu << a) & 0xFFFFu);
                const uint32_t vb = b >= 16 ? 0xFFFFu : (b <= 0 ? 0u : (1u << b) - 1u);
                valid[u] = va & vb;
            }
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            if (mode == 1) {
                uint32_t brace = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) brace |= bits4(eqmask(w[k], 0x7B7B7B7Bu) | eqmask(w[k], 0x7D7D7D7Du)) << (4 * k);
                special[u] = brace & valid[u];
                cnt[u] = __popc(valid[u]) + __popc(special[u]);
            } else {
                // '\' followed, inside the arena, by a brace that does not start the next string
                const uint32_t b0 = eqmask(w[0], 0x5C5C5C5Cu), b1 = eqmask(w[1], 0x5C5C5C5Cu), b2 = eqmask(w[2], 0x5C5C5C5Cu),
                               b3 = eqmask(w[3], 0x5C5C5C5Cu);
                special[u] = 0;
                if (b0 | b1 | b2 | b3) {  // most chunks of ordinary text hold no backslash at all
                    const uint32_t bslash = bits4(b0) | (bits4(b1) << 4) | (bits4(b2) << 8) | (bits4(b3) << 12);
                    uint32_t brace = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) brace |= bits4(eqmask(w[k], 0x7B7B7B7Bu) | eqmask(w[k], 0x7D7D7D7Du)) << (4 * k);
                    const uint32_t next_brace = (brace >> 1) | ((nb[u] == '{' || nb[u] == '}') ? 0x8000u : 0u);
                    const uint32_t next_valid = (valid[u] >> 1) | (b > 16 ? 0x8000u : 0u);
                    const uint32_t fixbits = have_fix ? (sm.fix[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu : 0u;
                    special[u] = bslash & next_brace & next_valid & valid[u] & ~fixbits;
                }
                cnt[u] = __popc(valid[u]) - __popc(special[u]);
            }
        }
    };

    PHASE_INIT();
    for (;;) {
        __syncthreads();  // the previous span's shared state is dead
        PHASE_MARK(0);
        const uint32_t span = ie_scan::acquire_tile(sm.scan, ws.tile_counter);
        PHASE_MARK(1);
        if (span >= spans) return;
        TRACE(span, 0);
        const uint64_t tile_lo = (uint64_t)span * span_tiles;
        const uint32_t nt = (uint32_t)min((uint64_t)span_tiles, tiles - tile_lo);

        // ---- pass 1: size of the span ----------------------------------------------------------------------
        uint32_t my_bytes = 0;
        for (uint32_t m = 0; m < nt; ++m) {
            const int64_t t0 = pa0 + (int64_t)(tile_lo + m) * TILE_BYTES;  // absolute position of the tile's byte 0 (may precede base0)
            load_tile(t0);
            const bool have_fix = build_fix(t0, tile_lo + m + 1 == tiles);
            classify(t0, have_fix);
#pragma unroll
            for (int u = 0; u < CPT; ++u) my_bytes += cnt[u];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_bytes += __shfl_xor_sync(0xFFFFFFFFu, my_bytes, d);
        if (lane == 0) sm.warp_tot[warp] = my_bytes;
        __syncthreads();
        uint64_t span_total = 0;
#pragma unroll
        for (int wv = 0; wv < NT / 32; ++wv) span_total += sm.warp_tot[wv];
        TRACE(span, 1);
        if (tid == 0) ie_scan::st_state(ws.tile_state + span, (span == 0 ? ie_scan::FLAG_INC : ie_scan::FLAG_AGG) | span_total);
        PHASE_MARK(2);
        uint64_t gbase = ie_scan::lookback_wide<NT, 2>(sm.scan, ws.tile_state, span, span_total);
        PHASE_MARK(3);
        TRACE(span, 2);
        if (gbase + span_total > out_cap) {
            // The caller's arena is too small: nothing is written for this span.  The LAST span still publishes the
            // total the output needs as out_offs[n], so that the caller can tell (out_offs[n] > out_capacity) and size
            // its arena; the other offsets are incomplete then.
            if (tid == 0) { *ws.overflow = 1u; if (tile_lo + nt == tiles) out_offs[n] = gbase + span_total; }
            continue;
        }

        // ---- pass 2: tile by tile, positions known ---------------------------------------------------------------
        for (uint32_t m = 0; m < nt; ++m) {
            const uint64_t tile = tile_lo + m;
            const bool last = tile + 1 == tiles;
            const int64_t t0 = pa0 + (int64_t)tile * TILE_BYTES;
            load_tile(t0);  // L2 hit: pass 1 read the same bytes a few microseconds ago
            const bool have_fix = build_fix(t0, last);
            if (use_list) { if (tid < 2) sm.bound[tid] = (tid == 1 && last) ? n + 1 : ws.tile_first[tile + tid]; }
            else if (!have_fix) find_bounds(t0, last);
            classify(t0, have_fix);
#pragma unroll
            for (int u = 0; u < CPT; ++u) sm.cmask[64 * warp + 32 * u + lane] = valid[u] | (special[u] << 16);
            // scan in chunk order: within a warp (u = 0 lanes, then u = 1 lanes), then across warps
            uint32_t inc[CPT];
#pragma unroll
            for (int u = 0; u < CPT; ++u) {
                inc[u] = cnt[u];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc[u], d);
                    if ((int)lane >= d) inc[u] += y;
                }
            }
            const uint32_t tot0 = __shfl_sync(0xFFFFFFFFu, inc[0], 31);
            if (lane == 31) sm.warp_tot[warp] = tot0 + inc[1];
            __syncthreads();
            uint32_t wbase = 0, tile_total = 0;
#pragma unroll
            for (int wv = 0; wv < NT / 32; ++wv) {
                const uint32_t x = sm.warp_tot[wv];
                if (wv < (int)warp) wbase += x;
                tile_total += x;
            }
            const uint32_t pre[CPT] = {wbase + inc[0] - cnt[0], wbase + tot0 + inc[1] - cnt[1]};
#pragma unroll
            for (int u = 0; u < CPT; ++u) sm.cpre[64 * warp + 32 * u + lane] = pre[u];
            uint8_t* gout = out + gbase;
            const uint32_t olead = (uint32_t)((uintptr_t)gout & 15);
            const bool staged = olead + tile_total <= (uint32_t)STAGE_CHUNKS * 16u;
            if (staged) {
                // the output bytes go to the staging buffer, laid out congruent to the output address modulo 16
#pragma unroll
                for (int u = 0; u < CPT; ++u) {
                    uint32_t o = olead + pre[u];
                    const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                    if (valid[u] == 0xFFFFu && special[u] == 0) {
                        // plain chunk: 16 bytes to one unaligned position = 3 shifted words + the 4 bytes around them
                        const uint32_t r = o & 3u, wi = o >> 2;
                        if (r == 0) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) sm.stage[(wi + k) + ((wi + k) >> 2)] = w[k];
                        } else {
                            const uint32_t sh = 8u * (4u - r);
#pragma unroll
                            for (int k = 0; k < 3; ++k) sm.stage[(wi + 1 + k) + ((wi + 1 + k) >> 2)] = __funnelshift_r(w[k], w[k + 1], sh);
#pragma unroll
                            for (int j = 0; j < 3; ++j) {  // head: bytes 0 .. 3-r of w[0]; tail: the top r bytes of w[3]
                                if (j < (int)(4u - r)) stage[stage_addr(o + j)] = (uint8_t)(w[0] >> (8 * j));
                                if (j < (int)r) stage[stage_addr(o + 16 - r + j)] = (uint8_t)(w[3] >> (8 * (4 - r + j)));
                            }
                        }
                    } else if (mode == 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (!((valid[u] >> j) & 1u)) continue;
                            if ((special[u] >> j) & 1u) stage[stage_addr(o++)] = '\\';
                            stage[stage_addr(o++)] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
                        }
                    } else {
                        const uint32_t keep = valid[u] & ~special[u];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if ((keep >> j) & 1u) stage[stage_addr(o++)] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
                    }
                }
            } else {  // brace-dense tile that does not fit the stage: per-thread byte stores
#pragma unroll
                for (int u = 0; u < CPT; ++u) {
                    uint8_t* d = gout + pre[u];
                    const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (!((valid[u] >> j) & 1u)) continue;
                        const bool sp = (special[u] >> j) & 1u;
                        if (mode == 1 && sp) *d++ = '\\';
                        if (mode == 1 || !sp) *d++ = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
                    }
                }
            }
            __syncthreads();  // stage, cpre, cmask, bound complete

            // out_offs of the strings that start in this tile
            const uint64_t i_lo = sm.bound[0], i_hi = sm.bound[1];
            for (uint64_t i = i_lo + tid; i < i_hi; i += NT) {
                const int64_t q = (int64_t)__ldg(offs + i) - t0;
                if (q >= TILE_BYTES && !last) continue;  // the closed end: owned by the next tile
                uint64_t o = gbase + tile_total;          // a start at end0 on the last tile
                if (q < TILE_BYTES) {
                    const uint32_t c = (uint32_t)q >> 4, below = (1u << (q & 15)) - 1u;
                    const uint32_t mk = sm.cmask[c];
                    const uint32_t va = mk & below, sp = (mk >> 16) & below;
                    o = gbase + sm.cpre[c] + (mode == 1 ? __popc(va) + __popc(sp) : __popc(va) - __popc(sp));
                }
                out_offs[i] = o;
            }
            // coalesced copy-out: full 16-byte chunks as one store, the ragged ends byte by byte
            if (staged && tile_total) {
                const uint32_t spn = olead + tile_total;
                const uint32_t o_chunks = (spn + 15) >> 4;
                uint8_t* g0 = gout - olead;  // 16-byte aligned
                for (uint32_t j = tid; j < o_chunks; j += NT) {
                    const uint32_t* sp = sm.stage + 5 * j;
                    const uint32_t lo = j == 0 ? olead : 0u, hi = min(16u, spn - 16 * j);
                    if (lo == 0 && hi == 16) *reinterpret_cast<uint4*>(g0 + 16 * (size_t)j) = make_uint4(sp[0], sp[1], sp[2], sp[3]);
                    else
                        for (uint32_t b = lo; b < hi; ++b) g0[16 * (size_t)j + b] = reinterpret_cast<const uint8_t*>(sp)[b];
                }
            }
            gbase += tile_total;
            __syncthreads();  // the next tile reuses stage / cpre / cmask / bound / warp_tot
        }
        PHASE_MARK(4);
        TRACE(span, 3);
    }
}

}  // namespace

cudaError_t ie_launch_escape(int mode, const uint8_t* d_in, const uint64_t* d_in_offs, uint64_t n, uint64_t in_bytes, uint8_t* d_out,
                             uint64_t out_cap, uint64_t* d_out_offs, const IeWorkspace& ws, cudaStream_t stream) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(ws.zero_base, 0, ws.zero_bytes, stream)) != cudaSuccess) return err;
    if (n == 0) return cudaMemsetAsync(d_out_offs, 0, sizeof(uint64_t), stream);
    const uint64_t max_tiles = ie_escape_tiles(in_bytes);
    ie_escape_prepass_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, stream>>>(mode, d_in, d_in_offs, n, ws.tile_first, max_tiles, ws.fix_count,
                                                                                 ws.fix_list);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent CTAs (span ids handed out in start order: the look-back never waits on an unscheduled CTA);
    // spans of up to 8 tiles, shorter when the arena is too small to give every resident CTA two spans
    const uint64_t resident = (uint64_t)sms * ESC_CTAS;
    uint64_t span_tiles = max_tiles / (2 * resident);
    span_tiles = span_tiles < 1 ? 1 : (span_tiles > 8 ? 8 : span_tiles);
    const uint64_t spans = (max_tiles + span_tiles - 1) / span_tiles;
    const uint64_t grid = spans < resident ? spans : resident;
    ie_escape_kernel<<<(unsigned)grid, NT, 0, stream>>>(mode, d_in, d_in_offs, n, d_out, out_cap, d_out_offs, ws, max_tiles, (uint32_t)span_tiles);
    return cudaGetLastError();
}
