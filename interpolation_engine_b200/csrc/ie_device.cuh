// Device helpers shared by the resolve kernels: table probe, the reference's leaf predicates, and the
// per-thread right-to-left traversal (the v1 fast path, kept as the exact fallback for tiles the
// cooperative tile kernel cannot hold in shared memory).
#pragma once
#include <cuda_runtime.h>

#include "ie_common.cuh"

namespace ie_dev {

constexpr uint32_t KCAP = 128;  // bytes of nested-key text a per-thread traversal can hold
constexpr uint32_t MAXLVL = 8;  // nesting depth of the fast paths

// ---- table lookup --------------------------------------------------------------------------
__device__ __forceinline__ const IeSlot* ie_lookup(const IeTableView& tv, const uint8_t* key, uint32_t len) {
    const uint32_t h = ie_hash_bytes(key, len);
    const IeSlot* slots = reinterpret_cast<const IeSlot*>(tv.base);
    uint32_t idx = h & tv.mask;
    for (;;) {
        const IeSlot* s = slots + idx;
        const uint4 hd = __ldg(reinterpret_cast<const uint4*>(s));  // hash, key_len, vl_tf, val_off16
        if (hd.y == IE_SLOT_EMPTY) return nullptr;
        if (hd.x == h && hd.y == len) {
            const uint8_t* stored = tv.base + (size_t)s->key_off16 * 16u;
            uint32_t i = 0;
            for (; i < len; ++i) if (__ldg(stored + i) != key[i]) break;
            if (i == len) return s;
        }
        idx = (idx + 1) & tv.mask;
    }
}

__device__ __forceinline__ bool is_arg_key(const uint8_t* k, uint32_t len) {  // interp.rs:109
    if (len < 3 || k[0] != 'A' || k[1] != 'R' || k[2] != 'G') return false;
    for (uint32_t i = 3; i < len; ++i) if (k[i] < '0' || k[i] > '9') return false;
    return true;
}
__device__ __forceinline__ bool tag_splices(uint32_t tag) {  // interp.rs:71-80
    return tag == IE_TAG_STRING || tag == IE_TAG_NUMBER || tag == IE_TAG_ARRAY;
}

// ---- fast path -------------------------------------------------------------------------------
struct Prescan {
    uint32_t n_open, n_close, m0;
    bool punt;
};

// One left-to-right pass: unescaped brace counts, the sentinel corner cases that the fast path
// does not reproduce, and m0 = min(leading '{' run, trailing unescaped '}' run).
__device__ __forceinline__ Prescan prescan(const uint8_t* __restrict__ t, uint32_t n) {
    Prescan ps{0, 0, 0, false};
    uint8_t p1 = 0, p2 = 0;  // previous two bytes
    uint32_t lead = 0;
    bool in_lead = true;
    for (uint32_t i = 0; i < n; ++i) {
        const uint8_t c = __ldg(t + i);
        if (c == '{') {
            if (p1 != '\\') { ++ps.n_open; if (in_lead) ++lead; }
            else in_lead = false;
        } else {
            in_lead = false;
            if (c == '}') {
                if (p1 != '\\') ++ps.n_close;
                else if (p2 == '.' || p2 == '}') ps.punt = true;  // ".\}" / "}\}": '.' + "〠." collides with ".〠"
            } else if (c == 0xA0 && p1 == 0x80 && p2 == 0xE3) ps.punt = true;  // literal U+3020
        }
        p2 = p1; p1 = c;
    }
    uint32_t trail = 0;
    for (uint32_t i = n; i > 0; --i) {
        if (__ldg(t + i - 1) != '}') break;
        if (i >= 2 && __ldg(t + i - 2) == '\\') break;
        ++trail;
    }
    ps.m0 = min(lead, trail);
    if (ps.n_open && ps.n_open != ps.n_close) ps.punt = true;  // uneven: the general path emits the exact text
    return ps;
}

template <bool WRITE>
__device__ __forceinline__ void fast_traverse(const IeTableView& tv, const uint8_t* __restrict__ t, uint32_t n, uint32_t m0,
                                              uint8_t* wend, uint32_t known_status, uint32_t& out_len, uint32_t& status,
                                              uint32_t& aux) {
    uint8_t kbuf[KCAP];
    uint32_t lvl_mark[MAXLVL], lvl_pos[MAXLVL];
    uint32_t ktop = KCAP, lvl = 0, p = n, olen = 0;
    uint8_t* w = wend;
    const bool emit = WRITE && known_status == IE_RES_STRING;
    status = IE_RES_STRING;
    aux = 0;
    while (p > 0) {
        const uint8_t c = __ldg(t + p - 1);
        const bool brace = (c == '{') || (c == '}');
        if (brace && p >= 2 && __ldg(t + p - 2) == '\\') {  // escaped brace: opaque pair
            if (lvl) {
                if (ktop < 2) { status = IE_RES_PUNT; break; }
                kbuf[--ktop] = c; kbuf[--ktop] = '\\';
            } else {
                olen += 2;
                if (emit) { *--w = c; *--w = '\\'; }
            }
            p -= 2;
            continue;
        }
        if (c == '}') {
            if (lvl == MAXLVL) { status = IE_RES_PUNT; break; }
            lvl_mark[lvl] = ktop; lvl_pos[lvl] = p - 1; ++lvl; --p;
            continue;
        }
        if (c == '{') {
            if (lvl == 0) { status = IE_RES_PANIC; olen = 0; break; }  // interp.rs:63-66
            --lvl;
            const uint32_t mark = lvl_mark[lvl], klen = mark - ktop, o = p - 1;
            const bool simple_layer = (o < m0) && (lvl_pos[lvl] == n - 1 - o);  // interp.rs:45-52
            const uint8_t* key = kbuf + ktop;
            uint32_t err = 0;
            const IeSlot* s = nullptr;
            if (klen == 0) err = IE_RES_EMPTY_KEY;
            else {
                s = ie_lookup(tv, key, klen);
                if (!s) err = is_arg_key(key, klen) ? IE_RES_ARG_MISSING : IE_RES_NOT_FOUND;
                else if (!simple_layer) {
                    const uint32_t tf = s->vl_tf;
                    if (!tag_splices(IE_SLOT_TAG(tf))) err = IE_RES_UNSUPPORTED;
                    else if (IE_SLOT_FLAGS(tf) & IE_VF_ANY) { status = IE_RES_PUNT; break; }
                }
            }
            if (err) {
                status = err; olen = klen;
                if (WRITE) for (uint32_t i = 0; i < klen; ++i) wend[(int)i - (int)klen] = key[i];
                break;
            }
            ktop = mark;
            const uint32_t vlen = IE_SLOT_VLEN(s->vl_tf);
            const uint8_t* v = tv.base + (size_t)s->val_off16 * 16u;
            if (lvl) {
                if (ktop < vlen) { status = IE_RES_PUNT; break; }
                ktop -= vlen;
                for (uint32_t i = 0; i < vlen; ++i) kbuf[ktop + i] = __ldg(v + i);
            } else if (simple_layer) {
                status = IE_RES_TYPED | (IE_SLOT_TAG(s->vl_tf) << 8);
                aux = s->entry;
                olen = vlen;
                if (WRITE) for (uint32_t i = 0; i < vlen; ++i) wend[(int)i - (int)vlen] = __ldg(v + i);
            } else {
                olen += vlen;
                if (emit) { w -= vlen; for (uint32_t i = 0; i < vlen; ++i) w[i] = __ldg(v + i); }
            }
            --p;
            continue;
        }
        if (lvl) {
            if (!ktop) { status = IE_RES_PUNT; break; }
            kbuf[--ktop] = c;
        } else {
            ++olen;
            if (emit) *--w = c;
        }
        --p;
    }
    out_len = (status == IE_RES_PUNT) ? 0u : olen;
}


}  // namespace ie_dev
