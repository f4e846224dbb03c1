// Leaf helpers shared by the two tile kernels of the batched resolver (ie_resolve_tile.cu: the phase-wise kernel with
// rescan rounds; ie_resolve_fused.cu: the single-pass kernel of the no-rounds launch): the SIMD-in-register chunk scan of
// P1, the unaligned 16-byte gathers of the lookups and of the copy sweep, and murmur3 over a key held in registers.
#pragma once
#include <cuda_runtime.h>

#include "ie_common.cuh"

namespace ie_tile {

#ifndef IE_P1_BATCH
#define IE_P1_BATCH 3
#endif
constexpr int P1_BATCH = IE_P1_BATCH;

__device__ __forceinline__ uint32_t eqmask(uint32_t w, uint32_t pat) {  // 0x80 in every byte of w equal to pat's
    const uint32_t x = w ^ pat;
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x) & 0x80808080u;
}
// Appends the 4 byte flags (0x80 per byte) of one word to a per-chunk mask: acc = acc << 4 | flags.  The multiply
// gathers bits 7 / 15 / 23 / 31 into the top nibble (no two partial products meet, so no carries).
__device__ __forceinline__ uint32_t push4(uint32_t acc, uint32_t flags) { return __funnelshift_l(flags * 0x00204081u, acc, 4); }

// The two rare corrections of scan_chunk, out of line: the chunk scan is unrolled P1_BATCH times and the kernel's
// instruction footprint matters.  `ec` = escaped '}' per byte (16 bits), `m` = the chunk's mask so far.
static __device__ __noinline__ uint32_t scan_chunk_rare(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t ec, int32_t p0, uint32_t tile_bytes,
                                                 const uint8_t* __restrict__ tp, uint32_t m) {
    const uint32_t w[4] = {w0, w1, w2, w3};
    // rare: escaped '}' preceded by '.' or '}' (".\}" / "}\}": '.' + "〠." reads as ".〠" + '.')
    while (ec) {
        const int j = __ffs(ec) - 1;
        ec &= ec - 1;
        const int32_t p = p0 + j;
        if (p >= 2) {
            const uint8_t b2 = __ldg(tp + p - 2);
            if (b2 == '.' || b2 == '}') m |= 0x10001u << j;
        }
    }
    // rare: literal U+3020 (E3 80 A0) collides with the reference's sentinels
    if ((w0 | w1 | w2 | w3) & 0x80808080u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t me = eqmask(w[k], 0xE3E3E3E3u);
            while (me) {
                const int bit = __ffs(me) - 1;
                me &= me - 1;
                const int32_t p = p0 + 4 * k + (bit >> 3);
                if (p >= 0 && (uint32_t)p + 2 < tile_bytes && __ldg(tp + p + 1) == 0x80 && __ldg(tp + p + 2) == 0xA0)
                    m |= 0x10001u << (4 * k + (bit >> 3));
            }
        }
    }
    return m;
}

// One 16-byte chunk of template text -> 32 bits: bit j = unescaped '{' at byte j, bit 16 + j = unescaped '}', both =
// punt marker.  `prev` is the byte before the chunk ("previous byte is a backslash" is evaluated on the flat stream;
// P2 repairs the first byte of each template).
// `escaped` (optional): set to the chunk's braces that the byte before them escapes (16 bits).
__device__ __forceinline__ uint32_t scan_chunk(const uint4& v, uint32_t prev, int32_t p0, uint32_t tile_bytes,
                                               const uint8_t* __restrict__ tp, uint32_t* escaped = nullptr) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t O = 0, C = 0, B = 0;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        O = push4(O, eqmask(w[k], 0x7B7B7B7Bu));
        C = push4(C, eqmask(w[k], 0x7D7D7D7Du));
        B = push4(B, eqmask(w[k], 0x5C5C5C5Cu));
    }
    if (p0 < 0 || p0 + 16 > (int32_t)tile_bytes) {  // first / last chunk: drop the bytes outside the tile
        const uint32_t lo = p0 < 0 ? (uint32_t)-p0 : 0u;
        const uint32_t hi = (int32_t)tile_bytes - p0 >= 16 ? 16u : (uint32_t)((int32_t)tile_bytes - p0);
        const uint32_t keep = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
        O &= keep; C &= keep; B &= keep;
    }
    const uint32_t esc = (B << 1) | (prev == '\\' ? 1u : 0u);
    uint32_t m = (O & ~esc) | ((C & ~esc) << 16);
    const uint32_t ec = C & esc;
    if (escaped) *escaped = (O | C) & esc;
    if (ec | ((w[0] | w[1] | w[2] | w[3]) & 0x80808080u)) m = scan_chunk_rare(w[0], w[1], w[2], w[3], ec, p0, tile_bytes, tp, m);
    return m;
}

// 16 bytes from an arbitrary address through ALIGNED 16-byte loads (the second one only when the m requested
// bytes reach into it) and register selects.  The kernel is bound by L1 wavefronts, not by ALU work: two
// vector loads per lane replace five scalar ones.  Bytes at index >= m are unspecified.
__device__ __forceinline__ uint4 load16_any(const uint8_t* __restrict__ p, uint32_t m) {
    const uintptr_t a = (uintptr_t)p;
    const uint4* ap = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint32_t r = (uint32_t)(a & 15);
    const uint4 A = __ldg(ap);
    uint4 B = make_uint4(0, 0, 0, 0);
    if (r + m > 16) B = __ldg(ap + 1);
    uint32_t w0 = A.x, w1 = A.y, w2 = A.z, w3 = A.w, w4 = B.x, w5 = B.y, w6 = B.z;
    if (r & 8) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = B.w; }
    if (r & 4) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
    const uint32_t sh = (r & 3) * 8;
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
// The 16 bytes at offset r (0..15) of the 32-byte quantity A | B << 128 (two consecutive aligned loads; B is not looked at
// when r is 0).
__device__ __forceinline__ uint4 align16(const uint4& A, const uint4& B, uint32_t r) {
    uint32_t w0 = A.x, w1 = A.y, w2 = A.z, w3 = A.w, w4 = B.x, w5 = B.y, w6 = B.z;
    if (r & 8) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = B.w; }
    if (r & 4) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
    const uint32_t sh = (r & 3) * 8;
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
// The 16-byte window at an arbitrary address p, restricted to its bytes [lo, hi) (0 <= lo < hi <= 16): byte j of the
// result is p[j] inside the range and zero outside.  Only the aligned 16-byte blocks that hold a requested byte are
// read, so p may be a VIRTUAL address: "where the piece would start if it began at byte 0 of the destination chunk".
// A piece lands at its place in a destination chunk without any register shift: acc |= load16_range(src - lo, lo, hi).
__device__ __forceinline__ uint4 load16_range(const uint4* __restrict__ lowmask, const uint8_t* __restrict__ p, uint32_t lo, uint32_t hi) {
    const uintptr_t a = (uintptr_t)p;
    const uint4* ap = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint32_t r = (uint32_t)(a & 15);
    uint4 A = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0, 0);
    if (lo + r < 16) A = __ldg(ap);
    if (hi + r > 16) B = __ldg(ap + 1);
    uint32_t w0 = A.x, w1 = A.y, w2 = A.z, w3 = A.w, w4 = B.x, w5 = B.y, w6 = B.z;
    if (r & 8) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = B.w; }
    if (r & 4) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
    const uint32_t sh = (r & 3) * 8;
    const uint4 mh = lowmask[hi], ml = lowmask[lo];
    return make_uint4(__funnelshift_r(w0, w1, sh) & mh.x & ~ml.x, __funnelshift_r(w1, w2, sh) & mh.y & ~ml.y,
                      __funnelshift_r(w2, w3, sh) & mh.z & ~ml.z, __funnelshift_r(w3, w4, sh) & mh.w & ~ml.w);
}
// acc |= v << (8 * s bytes), s in [0, 15], as one 128-bit little-endian quantity
__device__ __forceinline__ void or_shifted(uint4& acc, const uint4& v, uint32_t s) {
    const uint64_t lo = (uint64_t)v.x | ((uint64_t)v.y << 32), hi = (uint64_t)v.z | ((uint64_t)v.w << 32);
    uint64_t rlo, rhi;
    if (s == 0) { rlo = lo; rhi = hi; }
    else if (s < 8) { rlo = lo << (8 * s); rhi = (hi << (8 * s)) | (lo >> (64 - 8 * s)); }
    else if (s == 8) { rlo = 0; rhi = lo; }
    else { rlo = 0; rhi = lo << (8 * (s - 8)); }
    acc.x |= (uint32_t)rlo; acc.y |= (uint32_t)(rlo >> 32); acc.z |= (uint32_t)rhi; acc.w |= (uint32_t)(rhi >> 32);
}
__device__ __forceinline__ uint32_t hash_short(const uint4& k, uint32_t klen) {  // == ie_hash_bytes on the same bytes
    uint32_t h = 0x9747b28cu;
    const uint32_t nb = klen >> 2;
    if (nb > 0) h = ie_mur_step(h, k.x);
    if (nb > 1) h = ie_mur_step(h, k.y);
    if (nb > 2) h = ie_mur_step(h, k.z);
    if (nb > 3) h = ie_mur_step(h, k.w);
    if (klen & 3) h = ie_mur_tail(h, nb == 0 ? k.x : nb == 1 ? k.y : nb == 2 ? k.z : k.w);
    return ie_fmix32(h ^ klen);
}

}  // namespace ie_tile
