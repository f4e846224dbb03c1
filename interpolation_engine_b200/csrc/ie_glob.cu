// Wildcard delete sweep for sm_100a: wildcard_match (runtime.rs:1633-1647) of every key of an
// inserts map against the wildcard list of a `delete` / `delete_except` task (runtime.rs:1198-1239).
//
// The reference builds and compiles "^" + (".*" | literal)... + "$" with dot_matches_new_line for
// every (pattern, key) pair; the language of that regex is "literal runs separated by arbitrary
// gaps", matched here directly on bytes (equivalent on valid UTF-8).  One thread per key,
// 32 keys per warp -> one coalesced mask word per warp via ballot.  Patterns ride in the kernel
// parameter block (constant bank, broadcast reads).
#include <cuda_runtime.h>

#include "ie_kernels.h"

namespace {

__device__ __forceinline__ bool glob_match(const uint8_t* __restrict__ p, uint32_t pn, const uint8_t* __restrict__ s, uint32_t sn) {
    uint32_t pi = 0, si = 0, star = 0xFFFFFFFFu, mark = 0;
    while (si < sn) {
        if (pi < pn && p[pi] == '*') { star = pi++; mark = si; }
        else if (pi < pn && p[pi] == s[si]) { ++pi; ++si; }
        else if (star != 0xFFFFFFFFu) { pi = star + 1; si = ++mark; }
        else return false;
    }
    while (pi < pn && p[pi] == '*') ++pi;
    return pi == pn;
}

constexpr int GLOB_CTA = 256;           // keys per CTA
constexpr uint32_t STAGE_BYTES = 24576;  // key text of one CTA staged in shared memory (96 bytes per key on average)

// The 256 keys of a CTA are contiguous in the arena: their text is staged with coalesced 16-byte loads,
// then every thread matches its own key out of shared memory (tiles whose text exceeds the stage read
// global memory directly).
__global__ void __launch_bounds__(GLOB_CTA) ie_glob_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ offs, uint64_t n,
                                                           const __grid_constant__ IeGlobPatterns pats, uint32_t* __restrict__ mask,
                                                           unsigned long long* __restrict__ n_deleted) {
    __shared__ __align__(16) uint8_t stage[STAGE_BYTES + 32];
    __shared__ unsigned int s_deleted;
    const uint64_t k0 = (uint64_t)blockIdx.x * GLOB_CTA;
    const uint64_t k = k0 + threadIdx.x;
    const uint32_t nk = (uint32_t)min((uint64_t)GLOB_CTA, n - k0);
    const uint64_t b0 = __ldg(offs + k0), b1 = __ldg(offs + k0 + nk);
    const uint64_t a = k < n ? __ldg(offs + k) : b1;
    const uint64_t a_next = k < n ? __ldg(offs + k + 1) : b1;
    if (threadIdx.x == 0) s_deleted = 0;
    const uintptr_t g0 = (uintptr_t)(keys + b0) & ~(uintptr_t)15;   // aligned floor of the tile's first byte
    const uint32_t lead = (uint32_t)((uintptr_t)(keys + b0) - g0);
    const bool staged = (b1 - b0) + lead <= STAGE_BYTES;
    if (staged) {
        const uint32_t chunks = (uint32_t)((b1 - b0) + lead + 15) >> 4;
        for (uint32_t c = threadIdx.x; c < chunks; c += GLOB_CTA)
            *reinterpret_cast<uint4*>(stage + 16 * c) = __ldg(reinterpret_cast<const uint4*>(g0) + c);
    }
    __syncthreads();
    bool del = false;
    if (k < n) {
        const uint32_t len = (uint32_t)(a_next - a);
        const uint8_t* s = staged ? stage + lead + (uint32_t)(a - b0) : keys + a;
        bool any = false;
        for (uint32_t q = 0; q < pats.n_pat && !any; ++q)
            any = glob_match(pats.bytes + pats.off[q], (uint32_t)pats.off[q + 1] - pats.off[q], s, len);
        del = any != (pats.invert != 0);
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, del);
    if ((threadIdx.x & 31) == 0) {
        if (k < n) mask[k >> 5] = word;
        if (word) atomicAdd(&s_deleted, (unsigned int)__popc(word));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_deleted) atomicAdd(n_deleted, (unsigned long long)s_deleted);
}

}  // namespace

cudaError_t ie_launch_glob(const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n, const IeGlobPatterns& pats,
                           uint32_t* d_mask, uint64_t* d_n_deleted, cudaStream_t stream) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(d_n_deleted, 0, sizeof(uint64_t), stream)) != cudaSuccess) return err;
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + GLOB_CTA - 1) / GLOB_CTA;
    ie_glob_kernel<<<(unsigned)blocks, GLOB_CTA, 0, stream>>>(d_keys, d_key_offs, n, pats, d_mask,
                                                        reinterpret_cast<unsigned long long*>(d_n_deleted));
    return cudaGetLastError();
}
