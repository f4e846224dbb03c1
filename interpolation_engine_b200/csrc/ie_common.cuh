// Shared device/host definitions for the B200 interpolation engine.
//
// Device table layout (one allocation, `base`; all internal references are 16-byte units from
// `base`, so a table may span 64 GiB):
//
//   [ IeSlot slots[capacity] ][ key arena (keys > 16 B, 16 B aligned) ][ value arena (values > 16 B) ]
//
// A slot is 64 B = two 32 B sectors of one 128 B line.  Keys and values of <= 16 B live inside the
// slot itself (key_off16 / val_off16 then point at the slot's own inline area), so a C4-style
// lookup chain `{q-{idx-{slot-A}}}` costs one L2 round trip per hop instead of three.
#pragma once
#include <cstdint>

#include "../../include/ie_b200.h"

#define IE_SLOT_EMPTY 0xFFFFFFFFu
#define IE_INLINE_BYTES 16u

// value flags, computed once at pack time (ie_table.cpp: classify_value)
#define IE_VF_BRACE 1u     // contains an unescaped '{' or '}'  -> would be rescanned (interp.rs:81-83)
#define IE_VF_TRAIL_BS 2u  // ends with '\'                     -> may escape the brace that follows it
#define IE_VF_QUIRK 4u     // contains E3 80 A0 ("〠"), contains ".\}" or starts with "\}":
                           //   the sentinel encoding of interp.rs:40-43 is not injective there
#define IE_VF_ANY 7u
#define IE_VF_BALANCED 8u  // with IE_VF_BRACE and nothing else: the unescaped braces nest properly, so rescanning the
                           //   spliced value equals resolving the value's own groups in place (tile kernel: "splice")

// internal per-template status while a batch is in flight
#define IE_RES_PUNT 0xFF  // fast path declined; the general kernel resolves it

// A probe reads the slot's FIRST 32 bytes - everything a hit needs (hash, key length, value length / tag / flags, value
// reference) and the inline key - with ONE 256-bit load (LDG.E.256, sm_100a): one L1 tag request and one L2 round trip
// per probe.  The second 32 bytes (entry index, key reference, the inline value) are read only when the value itself is
// needed right away (the next hop of a `{q-{idx-{slot-A}}}` chain, a typed simple-path result).
struct __align__(32) IeSlot {
    uint32_t hash;       // murmur3_32(key)
    uint32_t key_len;    // IE_SLOT_EMPTY when free, IE_SLOT_TOMB when deleted
    uint32_t vl_tf;      // value length (25 bits) | tag << 25 | value flags << 28
    uint32_t val_off16;  // value bytes at base + 16 * val_off16
    uint8_t key_inline[IE_INLINE_BYTES];
    uint32_t entry;      // index of the insert in the caller's packed arrays (n, n+1: clock keys)
    uint32_t key_off16;  // key bytes at base + 16 * key_off16
    uint32_t pad[2];     // pad[0]: claim word of the device-side build
    uint8_t val_inline[IE_INLINE_BYTES];
};
#define IE_VLEN_MAX 0x01FFFFFFu
#define IE_SLOT_VLEN(x) ((x) & IE_VLEN_MAX)
#define IE_SLOT_TAG(x) (((x) >> 25) & 7u)
#define IE_SLOT_FLAGS(x) ((x) >> 28)
static_assert(sizeof(IeSlot) == 64, "slot must be 64 bytes");

struct IeTableView {
    const uint8_t* base;  // device pointer
    uint32_t mask;        // capacity - 1 (capacity is a power of two >= 2 * entries)
    uint32_t n_entries;
};

#if defined(__CUDACC__)
#define IE_HD __host__ __device__ __forceinline__
#else
#define IE_HD inline
#endif

IE_HD uint32_t ie_rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
IE_HD uint32_t ie_fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
IE_HD uint32_t ie_mur_step(uint32_t h, uint32_t k) {
    k *= 0xcc9e2d51u; k = ie_rotl32(k, 15); k *= 0x1b873593u;
    h ^= k; h = ie_rotl32(h, 13); h = h * 5u + 0xe6546b64u;
    return h;
}
IE_HD uint32_t ie_mur_tail(uint32_t h, uint32_t k) {  // k holds 1..3 trailing bytes, little endian
    k *= 0xcc9e2d51u; k = ie_rotl32(k, 15); k *= 0x1b873593u;
    return h ^ k;
}
// murmur3_32 over bytes (little-endian words), seed fixed.
IE_HD uint32_t ie_hash_bytes(const uint8_t* p, uint32_t len) {
    uint32_t h = 0x9747b28cu;
    uint32_t i = 0;
    for (; i + 4 <= len; i += 4) {
        uint32_t k = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
        h = ie_mur_step(h, k);
    }
    uint32_t k = 0;
    switch (len & 3u) {
        case 3: k |= (uint32_t)p[i + 2] << 16; /* fallthrough */
        case 2: k |= (uint32_t)p[i + 1] << 8;  /* fallthrough */
        case 1: k |= (uint32_t)p[i]; h = ie_mur_tail(h, k);
    }
    return ie_fmix32(h ^ len);
}

// Value flags (IE_VF_*), computed once per insert when a table is built or patched: what makes the fast path hand a
// template to the general path when this value is spliced into text (interp.rs:81-83 rescans the spliced value;
// interp.rs:40-43's sentinels collide).  Host (ie_table.cpp) and device (ie_table_build.cu) share this definition.
IE_HD uint32_t ie_classify_value(const uint8_t* v, uint64_t n) {
    uint32_t f = 0;
    int64_t depth = 0;
    bool nested_ok = true;
    if (n && v[n - 1] == '\\') f |= IE_VF_TRAIL_BS;
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t c = v[i];
        if (c == '{' || c == '}') {
            const bool esc = i > 0 && v[i - 1] == '\\';
            if (!esc) {
                f |= IE_VF_BRACE;
                depth += c == '{' ? 1 : -1;
                if (depth < 0) nested_ok = false;
            }
            else if (c == '}' && (i == 1 || v[i - 2] == '.' || v[i - 2] == '}')) f |= IE_VF_QUIRK;
        } else if (c == 0xA0 && i >= 2 && v[i - 1] == 0x80 && v[i - 2] == 0xE3) f |= IE_VF_QUIRK;
    }
    if (f == IE_VF_BRACE && nested_ok && depth == 0) f |= IE_VF_BALANCED;
    return f;
}

// A deleted slot (ie_table_delete): not empty, so probe chains run through it, and its key length matches no key.
#define IE_SLOT_TOMB 0xFFFFFFFEu
