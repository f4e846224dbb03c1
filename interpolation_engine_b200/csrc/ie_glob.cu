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

constexpr uint32_t KEY_REG_BYTES = 64;  // keys up to this size are staged in registers/local once

__global__ void __launch_bounds__(256) ie_glob_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ offs, uint64_t n,
                                                      const __grid_constant__ IeGlobPatterns pats, uint32_t* __restrict__ mask,
                                                      unsigned long long* __restrict__ n_deleted) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool del = false;
    if (k < n) {
        const uint64_t a = __ldg(offs + k);
        const uint32_t len = (uint32_t)(__ldg(offs + k + 1) - a);
        const uint8_t* key = keys + a;
        uint8_t local[KEY_REG_BYTES];
        const uint8_t* s = key;
        if (len <= KEY_REG_BYTES) {
            for (uint32_t i = 0; i < len; ++i) local[i] = __ldg(key + i);
            s = local;
        }
        bool any = false;
        for (uint32_t q = 0; q < pats.n_pat && !any; ++q)
            any = glob_match(pats.bytes + pats.off[q], (uint32_t)pats.off[q + 1] - pats.off[q], s, len);
        del = any != (pats.invert != 0);
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, del);
    if ((threadIdx.x & 31) == 0) {
        if (k < n) mask[k >> 5] = word;
        if (word) atomicAdd(n_deleted, (unsigned long long)__popc(word));
    }
}

}  // namespace

cudaError_t ie_launch_glob(const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n, const IeGlobPatterns& pats,
                           uint32_t* d_mask, uint64_t* d_n_deleted, cudaStream_t stream) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(d_n_deleted, 0, sizeof(uint64_t), stream)) != cudaSuccess) return err;
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + 255) / 256;
    ie_glob_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_keys, d_key_offs, n, pats, d_mask,
                                                        reinterpret_cast<unsigned long long*>(d_n_deleted));
    return cudaGetLastError();
}
