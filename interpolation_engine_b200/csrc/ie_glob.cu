// Wildcard delete sweep for sm_100a: wildcard_match (runtime.rs:1633-1647) of every key of an
// inserts map against the wildcard list of a `delete` / `delete_except` task (runtime.rs:1198-1239).
//
// The reference builds and compiles "^" + (".*" | literal)... + "$" with dot_matches_new_line for
// every (pattern, key) pair; the language of that regex is "literal runs separated by arbitrary
// gaps", matched here directly on bytes (equivalent on valid UTF-8).  One thread per key,
// 32 keys per warp -> one coalesced mask word per warp via ballot.  Patterns ride in the kernel
// parameter block (constant bank, broadcast reads).
//
// Patterns with at most one '*' (after collapsing runs of stars) and literal pieces of <= 32 bytes —
// every shape the reference's examples use ("persona/*", "*/field", "a*b", plain literals) — are compiled
// on the host into masked 32-byte prefix / suffix images: a key matches iff its length fits and its first
// and last 32 bytes agree under the masks, i.e. a handful of word compares on registers instead of a byte
// loop.  Anything else runs the generic byte matcher.
#include <cuda_runtime.h>

#include <cstring>

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

constexpr int GLOB_CTA = 256;            // threads per CTA
constexpr int MAX_KPT = 4;               // keys per thread and tile, picked from the mean key length
constexpr uint32_t STAGE_BYTES = 24576;  // key text of one tile staged in shared memory
constexpr uint32_t FRONT_PAD = 32, BACK_PAD = 48;  // the 32-byte windows of the first / last key may overhang the text
constexpr uint32_t BUF_BYTES = FRONT_PAD + STAGE_BYTES + BACK_PAD;
constexpr int GLOB_CTAS_PER_SM = 4;      // two stage buffers per CTA

// The staged key text is read through explicit shared-space loads: a pointer that may be shared OR global (the
// unstaged fallback) makes the compiler emit generic loads, which are slower than LDS.
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// 4 bytes at shared address a (any alignment)
__device__ __forceinline__ uint32_t lds_word(uint32_t a) {
    const uint32_t aw = a & ~3u;
    return __funnelshift_r(lds32(aw), lds32(aw + 4), (a & 3u) * 8);
}
// 32 bytes starting at shared address a (any alignment) as 8 little-endian words
__device__ __forceinline__ void load_window(uint32_t a, uint32_t (&w)[8]) {
    const uint32_t aw = a & ~3u, sh = (a & 3u) * 8;
    uint32_t x[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) x[k] = lds32(aw + 4 * k);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = __funnelshift_r(x[k], x[k + 1], sh);
}

// Returns 1 + the index of the first pattern that matches the key, 0 when none does.
__device__ __forceinline__ uint32_t match_key(const IeGlobPatterns& pats, const uint8_t* s, uint32_t len, bool staged) {
    // Per pattern, a PROBE first: the last word of the prefix image and the last word of the key against
    // the last word of the suffix image (the most discriminating ones: "persona-123/" differs from most
    // keys in "123/", "/field-42" in "d-42").  Only keys that pass the probe of a pattern with longer pieces
    // load their full 32-byte windows.
    const uint32_t sa = staged ? (uint32_t)__cvta_generic_to_shared(s) : 0u;  // the key's address in shared memory
    uint32_t t7 = 0;
    if (staged && pats.any_suf) t7 = lds_word(sa + len - 4);
    bool any = false;
    uint32_t q = 0;
    for (; q < pats.n_pat && !any; ++q) {
        const uint4 pa = *reinterpret_cast<const uint4*>(&pats.probe[q]);       // pre_word, pre_mask, suf_word, suf_mask
        const uint4 pb = *(reinterpret_cast<const uint4*>(&pats.probe[q]) + 1);  // min_len, max_len, probe_off, kind
        const uint32_t kind = pb.w & 0xFFu;
        if (staged && kind != IE_GLOB_GENERIC) {
            const uint32_t hw = lds_word(sa + pb.z);
            bool cand = (((hw ^ pa.x) & pa.y) | ((t7 ^ pa.z) & pa.w)) == 0 && len >= pb.x && len <= pb.y;
            if (cand && (pb.w >> 8) == 0) {  // pieces longer than the probe words: the full 32-byte windows
                const IeGlobFast& f = pats.fast[q];
                uint32_t H[8], T[8];  // the key's first / last 32 bytes (bytes outside the key are masked out)
                load_window(sa, H);
                load_window(sa + len - 32, T);
                uint32_t diff = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) diff |= ((H[w] ^ f.pre[w]) & f.pre_mask[w]) | ((T[w] ^ f.suf[w]) & f.suf_mask[w]);
                cand = diff == 0;
            }
            if (cand && kind == IE_GLOB_MID) {
                const IeGlobFast& f = pats.fast[q];
                if (len > 32) cand = glob_match(pats.bytes + pats.off[q], (uint32_t)pats.off[q + 1] - pats.off[q], s, len);
                else {
                    // Start positions of the middle piece: bytes of the key's first 32 equal to its FIRST byte (SIMD
                    // zero-byte test per word), limited to [mid_lo, len - mid_hi]; each candidate (a handful at most)
                    // then compares the piece's first word.
                    uint32_t H[8];
                    load_window(sa, H);
                    const uint32_t b0 = (f.mid & 0xFFu) * 0x01010101u;
                    uint32_t hits = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        const uint32_t x = H[w] ^ b0;
                        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
                        hits |= (((z >> 7) * 0x00204081u >> 21) & 0xFu) << (4 * w);
                    }
                    const uint32_t first = f.mid_lo, last = len - f.mid_hi;  // allowed start positions [first, last]
                    hits &= (last >= 31 ? 0xFFFFFFFFu : (2u << last) - 1u) & ~((1u << first) - 1u);
                    cand = false;
                    while (hits && !cand) {
                        const uint32_t i = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const uint32_t wv = lds_word(sa + i);
                        cand = ((wv ^ f.mid) & f.mid_mask) == 0;
                    }
                    if (cand && f.mid_len > 4) cand = glob_match(pats.bytes + pats.off[q], (uint32_t)pats.off[q + 1] - pats.off[q], s, len);
                }
            }
            any = cand;
        } else {
            any = glob_match(pats.bytes + pats.off[q], (uint32_t)pats.off[q + 1] - pats.off[q], s, len);
        }
    }
    return any ? q : 0u;  // q was incremented past the matching pattern
}

// Persistent CTAs, grid-stride over tiles of 256 * kpt consecutive keys.  The text of a tile is contiguous in
// the arena: it is brought into shared memory by 16-byte cp.async copies, double buffered, so the copies of
// tile i+1 (and the offset loads of tile i+2's bounds) are in flight while tile i is matched out of shared
// memory.  Tiles whose text exceeds the stage are matched from global memory by the generic matcher.
__global__ void __launch_bounds__(GLOB_CTA, GLOB_CTAS_PER_SM) ie_glob_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ offs,
                                                                             uint64_t n, const __grid_constant__ IeGlobPatterns pats,
                                                                             uint32_t* __restrict__ mask, unsigned long long* __restrict__ n_deleted,
                                                                             uint32_t* __restrict__ first) {
    extern __shared__ __align__(16) uint8_t stage_raw[];  // [2][BUF_BYTES]
    __shared__ unsigned int s_deleted;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) s_deleted = 0;
    // keys per thread: the largest of 1, 2, 4 whose expected tile text fits the stage with 25 % headroom
    const uint64_t total = __ldg(offs + n) - __ldg(offs);
    uint32_t kpt = MAX_KPT;
    while (kpt > 1 && (total / n + 1) * 5 / 4 * (GLOB_CTA * kpt) > STAGE_BYTES) kpt >>= 1;
    const uint32_t tile_keys = GLOB_CTA * kpt;
    const uint64_t n_tiles = (n + tile_keys - 1) / tile_keys;

    struct Bounds { uint64_t b0, b1; };
    auto bounds_of = [&](uint64_t tile) {
        const uint64_t k0 = tile * tile_keys;
        const uint64_t k1 = min(n, k0 + tile_keys);
        return Bounds{__ldg(offs + k0), __ldg(offs + k1)};
    };
    auto issue_copy = [&](const Bounds& bd, uint32_t buf) {  // returns nothing; oversized tiles are not staged
        const uintptr_t g0 = (uintptr_t)(keys + bd.b0) & ~(uintptr_t)15;
        const uint32_t lead = (uint32_t)((uintptr_t)(keys + bd.b0) - g0);
        if ((bd.b1 - bd.b0) + lead <= STAGE_BYTES) {
            const uint32_t chunks = (uint32_t)((bd.b1 - bd.b0) + lead + 15) >> 4;
            const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(stage_raw + buf * BUF_BYTES + FRONT_PAD);
            for (uint32_t c = tid; c < chunks; c += GLOB_CTA)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16 * c), "l"(g0 + 16 * (uintptr_t)c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint64_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    Bounds cur = bounds_of(tile);
    issue_copy(cur, 0);
    uint32_t buf = 0, my_deleted = 0;
    for (; tile < n_tiles; tile += gridDim.x) {
        const uint64_t next = tile + gridDim.x;
        const uint64_t k0 = tile * tile_keys;
        // this tile's key offsets and the next tile's bounds travel while the text copies are in flight
        uint64_t a[MAX_KPT], a_next[MAX_KPT];
#pragma unroll
        for (int j = 0; j < MAX_KPT; ++j) {
            const uint64_t k = k0 + (uint64_t)j * GLOB_CTA + tid;
            a[j] = a_next[j] = 0;
            if (j < (int)kpt && k < n) { a[j] = __ldg(offs + k); a_next[j] = __ldg(offs + k + 1); }
        }
        Bounds nxt = cur;
        if (next < n_tiles) nxt = bounds_of(next);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (next < n_tiles) issue_copy(nxt, buf ^ 1);
        const uint32_t lead = (uint32_t)((uintptr_t)(keys + cur.b0) & 15);
        const bool staged = (cur.b1 - cur.b0) + lead <= STAGE_BYTES;
        const uint8_t* text = stage_raw + buf * BUF_BYTES + FRONT_PAD + lead;
#pragma unroll
        for (int j = 0; j < MAX_KPT; ++j) {
            if (j >= (int)kpt) break;
            const uint64_t k = k0 + (uint64_t)j * GLOB_CTA + tid;
            if (k0 + (uint64_t)j * GLOB_CTA >= n) break;  // uniform: nothing left in this row
            bool del = false;
            if (k < n) {
                const uint32_t len = (uint32_t)(a_next[j] - a[j]);
                const uint8_t* s = staged ? text + (uint32_t)(a[j] - cur.b0) : keys + a[j];
                const uint32_t hit = match_key(pats, s, len, staged);
                if (first) first[k] = hit - 1u;  // 0xFFFFFFFF = no pattern matches (goto_map / replace_map pick the FIRST match)
                del = (hit != 0) != (pats.invert != 0);
            }
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, del);
            if ((tid & 31) == 0 && k < n) { mask[k >> 5] = word; my_deleted += __popc(word); }
        }
        cur = nxt;  // stage[buf] is overwritten only after the next iteration's barrier, which every thread reaches after this point
        buf ^= 1;
    }
    if (my_deleted) atomicAdd(&s_deleted, my_deleted);
    __syncthreads();
    if (tid == 0 && s_deleted) atomicAdd(n_deleted, (unsigned long long)s_deleted);
}

}  // namespace

// Host: fills pats->fast[] from pats->bytes / off (see the header comment).
void ie_glob_compile(IeGlobPatterns* pats) {
    pats->any_pre = pats->any_suf = 0;
    for (uint32_t q = 0; q < pats->n_pat; ++q) {
        IeGlobFast& f = pats->fast[q];
        std::memset(&f, 0, sizeof f);
        std::memset(&pats->probe[q], 0, sizeof pats->probe[q]);
        pats->probe[q].kind = IE_GLOB_GENERIC;
        f.kind = IE_GLOB_GENERIC;
        f.suf_first = 8;
        const uint8_t* p = pats->bytes + pats->off[q];
        const uint32_t pn = (uint32_t)pats->off[q + 1] - pats->off[q];
        // split at the star runs: at most two of them
        uint32_t run_lo[2] = {0, 0}, run_hi[2] = {0, 0}, runs = 0;  // [lo, hi) of each run of stars
        for (uint32_t i = 0; i < pn; ++i)
            if (p[i] == '*') {
                if (i == 0 || p[i - 1] != '*') { if (runs < 2) run_lo[runs] = i; ++runs; }
                if (runs <= 2) run_hi[runs - 1] = i + 1;
            }
        if (runs > 2) continue;
        const uint32_t pre_len = runs ? run_lo[0] : pn;
        const uint32_t suf_len = runs ? pn - run_hi[runs - 1] : 0;
        const uint32_t mid_len = runs == 2 ? run_lo[1] - run_hi[0] : 0;
        if (pre_len > 32 || suf_len > 32 || mid_len > 32) continue;
        f.kind = runs == 2 ? IE_GLOB_MID : IE_GLOB_FAST;
        f.exact = runs == 0;
        f.min_len = (uint16_t)(pre_len + mid_len + suf_len);
        uint8_t img[32], msk[32];
        std::memset(img, 0, 32); std::memset(msk, 0, 32);
        for (uint32_t j = 0; j < pre_len; ++j) { img[j] = p[j]; msk[j] = 0xFF; }
        std::memcpy(f.pre, img, 32); std::memcpy(f.pre_mask, msk, 32);
        f.pre_words = (uint8_t)((pre_len + 3) / 4);
        std::memset(img, 0, 32); std::memset(msk, 0, 32);
        for (uint32_t j = 0; j < suf_len; ++j) { img[32 - suf_len + j] = p[pn - suf_len + j]; msk[32 - suf_len + j] = 0xFF; }
        std::memcpy(f.suf, img, 32); std::memcpy(f.suf_mask, msk, 32);
        f.suf_first = (uint8_t)(8 - (suf_len + 3) / 4);
        f.probe = (uint8_t)(f.pre_words ? f.pre_words - 1 : 0);
        f.complete = pre_len <= 4 && suf_len <= 4;
        if (runs == 2) {
            std::memset(img, 0, 4); std::memset(msk, 0, 4);
            for (uint32_t j = 0; j < mid_len && j < 4; ++j) { img[j] = p[run_hi[0] + j]; msk[j] = 0xFF; }
            std::memcpy(&f.mid, img, 4); std::memcpy(&f.mid_mask, msk, 4);
            f.mid_len = mid_len;
            f.mid_lo = (uint8_t)pre_len;               // the middle piece may start at pre_len ...
            f.mid_hi = (uint8_t)(suf_len + mid_len);   // ... up to len - suf_len - mid_len
        }
        IeGlobProbe& pr = pats->probe[q];
        pr.pre_word = f.pre[f.probe]; pr.pre_mask = f.pre_mask[f.probe];
        pr.suf_word = f.suf[7]; pr.suf_mask = f.suf_mask[7];
        pr.min_len = f.min_len; pr.max_len = f.exact ? f.min_len : 0xFFFFFFFFu;
        pr.probe_off = 4u * f.probe;
        pr.kind = f.kind | ((uint32_t)f.complete << 8);
        if (pre_len) pats->any_pre = 1;
        if (suf_len) pats->any_suf = 1;
    }
}

// ---- a few long texts against a list of patterns: one CTA per text --------------------------------------------------
// replace_map / goto_map (runtime.rs:1649-1731, 1085-1133) test ONE text - an LLM answer of kilobytes - against a dozen
// patterns and want the first that matches.  One thread per key would walk those kilobytes alone, byte by byte and pattern
// after pattern; here the CTA works on one (text, pattern) pair together.  A pattern is literal pieces separated by star
// runs: the text must start with the first piece and end with the last one, and the middle pieces must be found in
// order, each at its LEFTMOST position behind the previous one (greedy placement decides the language of
// "^lit(.*)lit...(.*)lit$" exactly).  Prefix / suffix: strided compare + barrier vote; a middle piece: every thread tests
// four start positions per round, a shared atomicMin picks the leftmost hit.  Patterns come from global memory (no limit
// on their length or number); one of up to LONG_PAT_STAGE bytes is staged in shared memory first.
namespace {
constexpr int LONG_CTA = 256;
constexpr uint32_t LONG_PAT_STAGE = 4096;
constexpr uint32_t NONE = 0xFFFFFFFFu;

__device__ __forceinline__ bool cta_equal(const uint8_t* a, const uint8_t* b, uint32_t n) {
    bool ok = true;
    for (uint32_t i = threadIdx.x; i < n; i += LONG_CTA) ok &= a[i] == b[i];
    return __syncthreads_and(ok) != 0;
}

// leftmost p in [from, last] with key[p, p + ln) == lit, NONE if there is none; uniform arguments, uniform result
__device__ uint32_t cta_find(const uint8_t* key, const uint8_t* lit, uint32_t ln, uint32_t from, uint32_t last, uint32_t* best) {
    for (uint64_t base = from; base <= last; base += LONG_CTA * 4) {
        uint32_t mine = NONE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // ascending per thread: the first hit is this thread's leftmost
            const uint64_t p = base + (uint64_t)k * LONG_CTA + threadIdx.x;
            if (mine == NONE && p <= last) {
                uint32_t j = 0;
                while (j < ln && key[p + j] == lit[j]) ++j;
                if (j == ln) mine = (uint32_t)p;
            }
        }
        if (threadIdx.x == 0) *best = NONE;
        __syncthreads();
        if (mine != NONE) atomicMin(best, mine);
        __syncthreads();
        const uint32_t b = *best;
        __syncthreads();  // everyone has read it before the next round resets it
        if (b != NONE) return b;
    }
    return NONE;
}

__device__ bool cta_match(const uint8_t* pat, uint32_t pn, const uint8_t* key, uint32_t len, uint32_t* best) {
    uint32_t a = 0;
    while (a < pn && pat[a] != '*') ++a;                 // [0, a) = the piece before the first star
    if (a == pn) return len == pn && cta_equal(pat, key, pn);
    uint32_t b = pn;
    while (pat[b - 1] != '*') --b;                       // [b, pn) = the piece behind the last star
    const uint32_t suf = pn - b;
    if ((uint64_t)a + suf > len) return false;
    if (!cta_equal(pat, key, a)) return false;
    if (!cta_equal(pat + b, key + (len - suf), suf)) return false;
    uint32_t pos = a;
    const uint32_t limit = len - suf;                    // middle pieces live in [pos, limit)
    uint32_t i = a;
    while (i < b) {
        while (i < b && pat[i] == '*') ++i;
        uint32_t j = i;
        while (j < b && pat[j] != '*') ++j;
        const uint32_t ln = j - i;
        if (ln) {
            if ((uint64_t)pos + ln > limit) return false;
            const uint32_t p = cta_find(key, pat + i, ln, pos, limit - ln, best);
            if (p == NONE) return false;
            pos = p + ln;
        }
        i = j;
    }
    return true;
}

__global__ void __launch_bounds__(LONG_CTA) ie_glob_first_long_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ offs, uint64_t n,
                                                                      const uint8_t* __restrict__ pats, const uint64_t* __restrict__ pat_offs,
                                                                      uint32_t n_pat, uint32_t* __restrict__ first) {
    __shared__ uint32_t best;
    __shared__ uint8_t stage[LONG_PAT_STAGE];
    for (uint64_t k = blockIdx.x; k < n; k += gridDim.x) {
        const uint8_t* key = keys + offs[k];
        const uint32_t len = (uint32_t)(offs[k + 1] - offs[k]);
        uint32_t res = NONE;
        for (uint32_t q = 0; q < n_pat; ++q) {
            const uint8_t* pat = pats + pat_offs[q];
            const uint32_t pn = (uint32_t)(pat_offs[q + 1] - pat_offs[q]);
            if (pn <= LONG_PAT_STAGE) {
                __syncthreads();  // the previous pattern is no longer being read
                for (uint32_t i = threadIdx.x; i < pn; i += LONG_CTA) stage[i] = pat[i];
                __syncthreads();
                pat = stage;
            }
            if (cta_match(pat, pn, key, len, &best)) { res = q; break; }
        }
        if (threadIdx.x == 0) first[k] = res;
    }
}
}  // namespace

cudaError_t ie_launch_glob_first_long(const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n, const uint8_t* d_pats,
                                      const uint64_t* d_pat_offs, uint32_t n_pat, uint32_t* d_first, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    ie_glob_first_long_kernel<<<(unsigned)(n < 1184 ? n : 1184), LONG_CTA, 0, stream>>>(d_keys, d_key_offs, n, d_pats, d_pat_offs, n_pat, d_first);
    return cudaGetLastError();
}

cudaError_t ie_launch_glob(const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n, const IeGlobPatterns& pats,
                           uint32_t* d_mask, uint64_t* d_n_deleted, uint32_t* d_first, cudaStream_t stream) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(d_n_deleted, 0, sizeof(uint64_t), stream)) != cudaSuccess) return err;
    if (n == 0) return cudaSuccess;
    if ((err = cudaFuncSetAttribute(ie_glob_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * BUF_BYTES))) != cudaSuccess) return err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t min_tiles = (n + GLOB_CTA * MAX_KPT - 1) / (GLOB_CTA * MAX_KPT);  // the kernel may pick smaller tiles: more of them
    const uint64_t resident = (uint64_t)sms * GLOB_CTAS_PER_SM;
    const uint64_t blocks = min_tiles < resident ? min_tiles : resident;
    ie_glob_kernel<<<(unsigned)blocks, GLOB_CTA, 2 * BUF_BYTES, stream>>>(d_keys, d_key_offs, n, pats, d_mask,
                                                                          reinterpret_cast<unsigned long long*>(d_n_deleted), d_first);
    return cudaGetLastError();
}
