// Cooperative tile kernel of the batched `{key}` resolver (sm_100a) — the hot path.
//
// Replaces interpolate_inserts + get_interpdata (rust-project/src/interp.rs:31-137) for one tile of
// IE_RESOLVE_TILE consecutive templates per CTA.  The work is split into phases that keep a warp in
// ONE kind of work at a time (the thread-per-template v1 kernel ran at 5 of 32 active lanes):
//
//   P1  flat scan      all lanes stream the tile's bytes as 16-byte coalesced chunks; SIMD-in-register
//                      byte compares give one 32-bit brace mask per chunk (2 bits per byte), stored
//                      direct-mapped in shared memory — no atomics, no compaction, no search
//   P2  structure      one thread per template: enumerates the set bits of its chunks, bracket
//                      matching, simple-path layers (interp.rs:45-52), leaf groups -> ready queue
//   P3  lookups        one thread per READY GROUP, level by level: key assembled in registers
//                      (literal pieces + inline child values), murmur3, one L2 round trip per probe
//   P4  sizes          one thread per template -> CTA scan + decoupled look-back -> compacted offsets,
//                      plus a tile-wide table of copy segments (literal runs and values)
//   P5  flat copy      all lanes sweep the tile's output range in 16-byte aligned chunks, gathering
//                      each chunk from its segment(s): coalesced 16-byte stores
//
// Exactness: identical argument to ie_device.cuh's per-thread traversal — with every spliced value
// free of unescaped braces and sentinel corner cases the reference's rightmost-first rewriting equals
// bracket matching, and the reference's error is the failing group with the largest '{' position.
// Anything else (flagged values, uneven or improper braces, sentinel collisions, capacity overflow of
// the per-tile tables) is handed to the general kernel or to the per-thread fallback, both exact.
//
// This file is compiled TWICE: as is (128-template tiles, 256 threads, 5 CTAs per SM) and with IE_TILE_SMALL
// (32-template tiles, 64 threads: many snapshots x a handful of templates each — the cloned-states shape, where a
// tile is one snapshot's templates and a 128-template tile would be a quarter full).
#ifdef IE_TILE_SMALL
#define IE_RESOLVE_TILE 32
#define IE_TILE_NT 64
#define IE_TILE_CTAS 20
#define ie_launch_resolve_tiles ie_launch_resolve_tiles_small
#endif
#include <cuda_runtime.h>

#include "ie_common.cuh"
#include "ie_device.cuh"
#include "ie_kernels.h"
#include "ie_scan.cuh"
#include "ie_tile_common.cuh"

#if defined(IE_PHASE_TIMING) && !defined(IE_TILE_SMALL)
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_phase_cycles[k], (unsigned long long)(now_ - t_phase_)); t_phase_ = now_; } } while (0)
#define PHASE_INIT() long long t_phase_ = clock64()
#else
#define PHASE_MARK(k) do { } while (0)
#define PHASE_INIT() do { } while (0)
#endif

// -DIE_DEBUG_BOUNDS: every index into a tile table is checked against the table's capacity before it is used; a
// violation is counted (and the first one recorded) instead of corrupting shared memory.  compute-sanitizer is not
// available on the GPU pool, this build takes its place: tests/fuzz_campaign.py runs against it unchanged and
// ie_debug_bound_violations() must stay 0 (profiles/r02_debug_bounds_fuzz.txt).
#ifdef IE_DEBUG_BOUNDS
namespace {
__device__ unsigned long long g_bound_violations[4];  // count, line, index, capacity of the first one (per translation unit)
__device__ __noinline__ void ie_bound_report(uint32_t i, uint32_t cap, int line) {
    if (atomicAdd(&g_bound_violations[0], 1ull) == 0) { g_bound_violations[1] = (unsigned long long)line; g_bound_violations[2] = i; g_bound_violations[3] = cap; }
}
__device__ __forceinline__ uint32_t ie_bound_check(uint32_t i, uint32_t cap, int line) {
    if (i >= cap) { ie_bound_report(i, cap, line); return 0u; }
    return i;
}
}  // namespace
#define IE_BOUND(i, cap) ie_bound_check((uint32_t)(i), (uint32_t)(cap), __LINE__)
#else
#define IE_BOUND(i, cap) (i)
#endif

namespace {

using namespace ie_dev;
using namespace ie_tile;

constexpr int TT = IE_RESOLVE_TILE;  // templates per tile at most (the launch picks tt <= TT from the mean template length)
#ifndef IE_TILE_NT
#define IE_TILE_NT (2 * IE_RESOLVE_TILE)
#endif
#ifndef IE_TILE_CTAS
#define IE_TILE_CTAS (640 / IE_RESOLVE_TILE)
#endif
constexpr int NT = IE_TILE_NT;       // threads per CTA
constexpr int NW = NT / 32;
constexpr int CTAS_PER_SM = IE_TILE_CTAS;  // resident CTAs the register budget is tuned for (48 registers)
#ifndef IE_E_PER2
#define IE_E_PER2 23  // brace events per template x 2 a tile's event arrays hold (ie_kernels.h: IE_TILE_EVENTS)
#endif
#ifndef IE_S_PER
#define IE_S_PER 8
#endif
constexpr int E_CAP = IE_E_PER2 * TT / 2;  // brace events per tile
constexpr int Q_CAP = E_CAP / 2;     // groups per tile
constexpr int E_PAD = E_CAP + E_CAP / 32 + 2;
constexpr int M_CAP = IE_M_PER * TT;  // 16-byte chunks per tile (288 bytes of template text per template)
constexpr int S_CAP = IE_S_PER * TT;  // copy segments per tile
#ifndef IE_C_PER
#define IE_C_PER 18
#endif
constexpr int C_CAP = IE_C_PER * TT;       // 16-byte output chunks with a segment index (288 bytes of output per template)
constexpr uint32_t POS_MASK = 0x00FFFFFFu;
constexpr uint32_t EV_SIMPLE = 0x80000000u;
constexpr uint32_t EV_CLOSE = 1u << 24;
constexpr uint32_t EV_DONE = 1u << 26;   // open event: the group is resolved, ev_a[open] / ev_a[close] hold its value
constexpr uint32_t EV_PUNT = 1u << 25;   // a byte the tile kernel does not interpret (sentinel collisions): the template is punted
constexpr uint32_t NONE16 = 0xFFFFu;
enum : uint32_t { TF_PUNT = 1, TF_VERBATIM = 2, TF_AGAIN = 4 };
constexpr uint32_t CS_EDGE = 0x8000u;    // cs[]: the chunk is not covered by ONE segment (pass B assembles it)

__device__ __forceinline__ uint32_t EI(uint32_t e) { return IE_BOUND(e + (e >> 5), E_PAD); }  // padded event index

struct Smem {
    ie_scan::TileSmemT<NT> scan;
    // Event arrays are indexed through EI(): one pad slot per 32 events, so that lanes walking their own
    // templates' events (about 8 apart) fall into different banks.
    uint32_t ev_pos[E_PAD];    // position in tile | EV_CLOSE (| EV_SIMPLE on opens)
    uint32_t ev_a[E_PAD];      // open: unresolved children, then val_off16 of the resolved value; close: its length
    uint16_t ev_match[E_PAD];  // partner event
    uint16_t ev_c[E_PAD];      // open: parent open (NONE16 = top level)
    union {
        struct {
            uint32_t cm[M_CAP];  // P1/P2: per chunk, bit j = unescaped '{' at byte j, bit 16 + j = '}', both = punt marker
            uint32_t q[Q_CAP];   // P2/P3: leaf groups (template << 16 | open event)
        } scan;
        struct {
            uint32_t out[S_CAP + 2];  // P5: tile-local output offset of each segment (+ sentinel)
            uint64_t src[S_CAP];      //     its source address
            uint16_t cs[C_CAP + 2];   //     segment holding the first byte of each 16-byte aligned output chunk | CS_EDGE
        } seg;
    } u;
    uint32_t t_start[TT + 1];  // template start, tile-relative
    uint32_t t_err[TT];        // max over failing groups of (open event << 8 | IE_RES_*)
    uint32_t t_splice[TT];     // 1 + the rightmost group whose value holds (balanced) groups of its own: another round
    uint32_t t_aux[TT];        // entry index of a typed (simple path) result
    uint32_t t_flags[TT];
    uint16_t t_eb[TT];         // first event of the template
    uint16_t t_ne[TT];         // its events
    uint8_t t_tag[TT];
    uint8_t t_layers[TT];      // simple-path layers of the template (interp.rs:45-52)
    uint4 lowmask[17];         // lowmask[k] = the low k bytes of a 16-byte quantity set (load16_range)
    uint32_t warp_scan[NW];
    uint32_t q_n[1];
    uint32_t q_lvl[3];         // P3: groups appended per level (rotating)
    uint32_t q_nb;             // leaf groups queued from the back of q[] (top-level leaves)
    uint32_t ev_n;             // events allocated
    uint32_t overflow;
};

// Five CTAs per SM must fit the 196 KB shared-memory carve-out step (each CTA also reserves 1 KB): one step further
// (228 KB) leaves the SM 28 KB of L1 instead of 60 KB, which costs this kernel 10 % (measured: 0.378 -> 0.415 ms).
static_assert(IE_RESOLVE_TILE != 128 || CTAS_PER_SM != 5 || 5 * (sizeof(Smem) + 1024) <= 196 * 1024, "Smem outgrew the 196 KB carve-out");


// Iterates the bytes of group g's key: literal template bytes and the values of its (resolved) children.
template <class F>
__device__ __forceinline__ void walk_key(const Smem& sm, const IeTableView& tv, const uint8_t* __restrict__ tp, uint32_t g, F& f) {
    uint32_t pos = (sm.ev_pos[EI(g)] & POS_MASK) + 1;
    const uint32_t c = sm.ev_match[EI(g)];
    uint32_t e = g + 1;
    for (;;) {
        const uint32_t stop = sm.ev_pos[EI(e)] & POS_MASK;
        for (; pos < stop; ++pos) if (!f(__ldg(tp + pos))) return;
        if (e == c) return;
        const uint32_t ce = sm.ev_match[EI(e)];
        const uint8_t* v = tv.base + (size_t)sm.ev_a[EI(e)] * 16u;
        const uint32_t vl = sm.ev_a[EI(ce)];
        for (uint32_t k = 0; k < vl; ++k) if (!f(__ldg(v + k))) return;
        pos = (sm.ev_pos[EI(ce)] & POS_MASK) + 1;
        e = ce + 1;
    }
}
struct Hasher {
    uint32_t h = 0x9747b28cu, w = 0, n = 0;
    __device__ __forceinline__ bool operator()(uint8_t b) {
        w |= (uint32_t)b << (8 * (n & 3));
        if ((++n & 3) == 0) { h = ie_mur_step(h, w); w = 0; }
        return true;
    }
    __device__ __forceinline__ uint32_t finish() const { return ie_fmix32(((n & 3) ? ie_mur_tail(h, w) : h) ^ n); }
};
struct Comparer {
    const uint8_t* s;
    uint32_t i = 0;
    bool ok = true;
    __device__ __forceinline__ bool operator()(uint8_t b) {
        if (__ldg(s + i) != b) { ok = false; return false; }
        ++i;
        return true;
    }
};
struct ArgCheck {  // interp.rs:109: "ARG" followed by ASCII digits only
    uint32_t i = 0;
    bool ok = true;
    __device__ __forceinline__ bool operator()(uint8_t b) {
        const bool good = i == 0 ? b == 'A' : i == 1 ? b == 'R' : i == 2 ? b == 'G' : (b >= '0' && b <= '9');
        ++i;
        if (!good) ok = false;
        return good;
    }
};

// Calls f(src, len) for each non-empty piece of group g's key (error payloads).
template <class F>
__device__ __forceinline__ void walk_key_pieces(const Smem& sm, const IeTableView& tv, const uint8_t* tp, uint32_t g, F& f) {
    uint32_t pos = (sm.ev_pos[EI(g)] & POS_MASK) + 1;
    const uint32_t c = sm.ev_match[EI(g)];
    uint32_t e = g + 1;
    for (;;) {
        const uint32_t stop = sm.ev_pos[EI(e)] & POS_MASK;
        if (stop > pos) f(tp + pos, stop - pos);
        if (e == c) return;
        const uint32_t ce = sm.ev_match[EI(e)];
        if (sm.ev_a[EI(ce)]) f(tv.base + (size_t)sm.ev_a[EI(e)] * 16u, sm.ev_a[EI(ce)]);
        pos = (sm.ev_pos[EI(ce)] & POS_MASK) + 1;
        e = ce + 1;
    }
}
// Calls f(src, len) for each non-empty output piece of a template: its text with every RESOLVED group (outermost
// first) replaced by the group's value.  ROUNDS: unresolved groups (only in templates that go another round) stay
// as text, with their own resolved descendants replaced; without rounds every top-level group is resolved here.
template <bool ROUNDS, class F>
__device__ __forceinline__ void walk_output_pieces(const Smem& sm, const IeTableView& tv, const uint8_t* tp, uint32_t t, F& f) {
    uint32_t pos = sm.t_start[t];
    const uint32_t end = sm.t_start[t + 1];
    uint32_t e = sm.t_eb[t];
    const uint32_t ee = e + sm.t_ne[t];
    while (e < ee) {
        const uint32_t v = sm.ev_pos[EI(e)];
        if (ROUNDS && (v & (EV_CLOSE | EV_DONE)) != EV_DONE) { ++e; continue; }  // a close, or an open that stays text
        const uint32_t o = v & POS_MASK;
        if (o > pos) f(tp + pos, o - pos);
        const uint32_t ce = sm.ev_match[EI(e)];
        if (sm.ev_a[EI(ce)]) f(tv.base + (size_t)sm.ev_a[EI(e)] * 16u, sm.ev_a[EI(ce)]);
        pos = (sm.ev_pos[EI(ce)] & POS_MASK) + 1;
        e = ce + 1;
    }
    if (end > pos) f(tp + pos, end - pos);
}
struct PieceCount {
    uint32_t bytes = 0, n = 0;
    __device__ __forceinline__ void operator()(const uint8_t*, uint32_t len) { bytes += len; ++n; }
};
struct PieceEmit {
    Smem& sm;
    uint32_t idx, off;
    uint32_t olead;  // bytes between the 16-byte aligned floor of the tile's output address and that address
    bool index_chunks;
    __device__ __forceinline__ void operator()(const uint8_t* src, uint32_t len) {
        sm.u.seg.out[IE_BOUND(idx, S_CAP)] = off;
        sm.u.seg.src[IE_BOUND(idx, S_CAP)] = (uint64_t)(uintptr_t)src;
        if (index_chunks) {
            // every 16-byte aligned output chunk whose first byte lies in this piece points back at it; only the last
            // of them can reach beyond the piece's end (CS_EDGE: pass B assembles that chunk)
            const uint32_t lo = off + olead, hi = lo + len;  // the piece in chunk coordinates
            uint32_t c = (lo + 15) >> 4;
            for (; (c << 4) + 16 <= hi; ++c) sm.u.seg.cs[IE_BOUND(c, C_CAP + 2)] = (uint16_t)idx;
            if ((c << 4) < hi) sm.u.seg.cs[IE_BOUND(c, C_CAP + 2)] = (uint16_t)(idx | CS_EDGE);
        }
        ++idx;
        off += len;
    }
};
struct PieceCopy {  // warp-wide byte copy (segment-table overflow fallback): all 32 lanes walk the same template
    uint8_t* dst;
    uint32_t lane;
    __device__ __forceinline__ void operator()(const uint8_t* src, uint32_t len) {
        for (uint32_t k = lane; k < len; k += 32) dst[k] = __ldg(src + k);
        dst += len;
    }
};

// Assembles group g's key into 16 bytes of registers (literal pieces + inline child values).
// Returns false when the key is longer than 16 bytes (the byte-walking path handles those).
// `carry_e` / `carry_v`: a child whose inline value the caller already holds in registers (the group it
// resolved just before walking up to g), saving the dependent reload of the slot.
__device__ __forceinline__ bool short_key(const Smem& sm, const IeTableView& tv, const uint8_t* __restrict__ tp, uint32_t g, uint4& key,
                                          uint32_t& klen, uint32_t carry_e, const uint4& carry_v) {
    key = make_uint4(0, 0, 0, 0);
    klen = 0;
    uint32_t pos = (sm.ev_pos[EI(g)] & POS_MASK) + 1;
    const uint32_t c = sm.ev_match[EI(g)];
    uint32_t e = g + 1;
    for (;;) {
        const uint32_t stop = sm.ev_pos[EI(e)] & POS_MASK;
        const uint32_t m = stop - pos;
        if (klen + m > 16) return false;
        if (m) {
            const uint4 v = load16_range(sm.lowmask, tp + pos - klen, klen, klen + m);
            key.x |= v.x; key.y |= v.y; key.z |= v.z; key.w |= v.w;
            klen += m;
        }
        if (e == c) return true;
        const uint32_t ce = sm.ev_match[EI(e)];
        const uint32_t vl = sm.ev_a[EI(ce)];
        if (klen + vl > 16) return false;
        if (vl) {  // values of <= 16 bytes live zero-padded in their slot's 16-byte aligned inline area
            if (e == carry_e) or_shifted(key, carry_v, klen);
            else {
                const uint4 v = load16_range(sm.lowmask, tv.base + (size_t)sm.ev_a[EI(e)] * 16u - klen, klen, klen + vl);
                key.x |= v.x; key.y |= v.y; key.z |= v.z; key.w |= v.w;
            }
            klen += vl;
        }
        pos = (sm.ev_pos[EI(ce)] & POS_MASK) + 1;
        e = ce + 1;
    }
}

// Resolves group g and then, while g was the last unresolved child of its parent, the parent too
// (a `{q-{idx-{slot-A}}}` chain is one thread's work instead of one queue round per level).
// IE_P3_LEVELS (experiment switch, profiles/r02_kernel_experiments.md): instead of carrying on with the parent itself the
// thread returns the parent that became ready (NONE16: none) and the caller queues it for the next level, so that a
// chain hop runs on a compacted queue at full lanes.  Measured on C4: 8 % fewer warp instructions (223 M vs 242 M,
// 27 instead of 25 active lanes) but 6 % MORE time (0.366 vs 0.343 ms): the later levels hold half as many groups as
// the CTA has threads, and the barrier per level costs more than the idle lanes of a chain.
template <bool ROUNDS>
__device__ __forceinline__ uint32_t resolve_group(Smem& sm, const IeTableView& tv, const uint8_t* __restrict__ tp, uint32_t t, uint32_t g) {
  uint32_t carry_e = NONE16;
  uint4 carry_v = make_uint4(0, 0, 0, 0);
  for (;;) {
    const uint32_t c = sm.ev_match[EI(g)];
    const bool simple = (sm.ev_pos[EI(g)] & EV_SIMPLE) != 0;
    uint32_t err = 0, klen, val_off16 = 0, vl_tf = 0;
    const IeSlot* hit = nullptr;
    uint4 key;
    const IeSlot* slots = reinterpret_cast<const IeSlot*>(tv.base);
    // the typed simple path reports the entry, a group inside another group hands its (inline) value to the parent's key
    const bool want_tail = simple || sm.ev_c[EI(g)] != NONE16;
    uint4 tail_hdr = make_uint4(0, 0, 0, 0), tail_val = make_uint4(0, 0, 0, 0);
    if (short_key(sm, tv, tp, g, key, klen, carry_e, carry_v)) {
        if (klen == 0) err = IE_RES_EMPTY_KEY;  // interp.rs:105
        else {
            const uint32_t h = hash_short(key, klen);
            uint32_t idx = h & tv.mask;
            for (;;) {
                // header and inline key come with ONE 256-bit load: one L1 tag request, one L2 round trip per probe; the
                // slot's second half (entry index, inline value) rides along only when a hit will need it right away
                uint4 q0, q2;
                asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w)
                             : "l"(slots + idx));
                if (want_tail)
                    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(tail_hdr.x), "=r"(tail_hdr.y), "=r"(tail_hdr.z), "=r"(tail_hdr.w), "=r"(tail_val.x), "=r"(tail_val.y),
                                   "=r"(tail_val.z), "=r"(tail_val.w)
                                 : "l"(reinterpret_cast<const uint8_t*>(slots + idx) + 32));
                if (q0.y == IE_SLOT_EMPTY) break;
                if (q0.x == h && q0.y == klen && q2.x == key.x && q2.y == key.y && q2.z == key.z && q2.w == key.w) {
                    hit = slots + idx; vl_tf = q0.z; val_off16 = q0.w;
                    break;
                }
                idx = (idx + 1) & tv.mask;
            }
            if (!hit) {  // interp.rs:109-116, :136
                bool arg = klen >= 3 && (key.x & 0x00FFFFFFu) == 0x00475241u;  // "ARG"
                for (uint32_t j = 3; arg && j < klen; ++j) {
                    const uint32_t wj = j < 4 ? key.x : j < 8 ? key.y : j < 12 ? key.z : key.w;
                    const uint32_t b = (wj >> (8 * (j & 3))) & 0xFFu;
                    arg = b >= '0' && b <= '9';
                }
                err = arg ? IE_RES_ARG_MISSING : IE_RES_NOT_FOUND;
            }
        }
    } else {
        Hasher hs;
        walk_key(sm, tv, tp, g, hs);
        klen = hs.n;
        const uint32_t h = hs.finish();
        uint32_t idx = h & tv.mask;
        for (;;) {
            const IeSlot* cand = slots + idx;
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(cand));
            if (q0.y == IE_SLOT_EMPTY) break;
            if (q0.x == h && q0.y == klen) {
                Comparer cmp{tv.base + (size_t)__ldg(&cand->key_off16) * 16u};
                walk_key(sm, tv, tp, g, cmp);
                if (cmp.ok) { hit = cand; vl_tf = q0.z; val_off16 = q0.w; break; }
            }
            idx = (idx + 1) & tv.mask;
        }
        if (!hit) {
            ArgCheck ac;
            walk_key(sm, tv, tp, g, ac);
            err = ac.ok ? IE_RES_ARG_MISSING : IE_RES_NOT_FOUND;
        }
    }
    bool splice = false;
    if (hit && !simple) {
        if (!tag_splices(IE_SLOT_TAG(vl_tf))) err = IE_RES_UNSUPPORTED;  // interp.rs:71-80
        else if (IE_SLOT_FLAGS(vl_tf) & IE_VF_ANY) {
            // interp.rs:81-83 rescans the spliced value.  Properly nested groups in it resolve in place: the value's
            // text takes the group's place and the template goes another round; anything else is the general path's.
            if (ROUNDS && IE_SLOT_FLAGS(vl_tf) == (IE_VF_BRACE | IE_VF_BALANCED)) splice = true;
            else { atomicOr(&sm.t_flags[t], TF_PUNT); return NONE16; }
        }
    }
    if (err) { atomicMax(&sm.t_err[t], (g << 8) | err); return NONE16; }
    sm.ev_a[EI(g)] = val_off16;
    sm.ev_a[EI(c)] = IE_SLOT_VLEN(vl_tf);
    if (ROUNDS) sm.ev_pos[EI(g)] |= EV_DONE;
    if (ROUNDS && splice) {  // the groups enclosing g cannot be looked up yet: they stay text for the next round
        atomicMax(&sm.t_splice[t], g + 1u);
        atomicOr(&sm.t_flags[t], TF_AGAIN);
        return NONE16;
    }
    const uint32_t parent = sm.ev_c[EI(g)];
    if (parent == NONE16) {
        if (simple) { sm.t_aux[t] = klen <= 16 ? tail_hdr.x : __ldg(&hit->entry); sm.t_tag[t] = (uint8_t)IE_SLOT_TAG(vl_tf); }
        return NONE16;
    }
    // one child of `parent` resolved; whoever resolves the last one carries on with the parent / hands it to the next level
    __threadfence_block();
    if (atomicSub(&sm.ev_a[EI(parent)], 1u) != 1u) return NONE16;
#ifdef IE_P3_LEVELS
    return parent;  // (the level barrier orders the siblings' results before the parent's lookup)
#endif
    __threadfence_block();  // the siblings' results (written before their decrements) are visible from here on
    // short-key hits leave the slot's inline value in tail_val (valid when the value is inline)
    carry_e = (klen <= 16 && IE_SLOT_VLEN(vl_tf) <= IE_INLINE_BYTES) ? g : NONE16;
    carry_v = tail_val;
    g = parent;
  }
}

// The cold blocks of the tile body, out of line (arguments by value: taking the address of a kernel-level struct would
// move it to local memory for the hot path too).  Measured: 0.379 -> 0.374 ms on C4 (the hot instructions are a
// sixth of the kernel's code and were spread over all of it).
// Segment-table overflow (a tile with more copy segments than S_CAP - one template with hundreds of groups is enough):
// the templates are copied one per WARP, every lane walking the same pieces and taking every 32nd byte.  (One thread
// per template took 11 ms for a single 16 KB template with 600 groups.)  The per-template parameters come from the
// dead segment table: five words per template.
template <bool ROUNDS>
__device__ __noinline__ void copy_own_pieces(Smem* smp, IeTableView tv0, const uint8_t* __restrict__ tp, uint32_t nt, uint8_t* out,
                                             const uint32_t* __restrict__ map_i0, uint32_t per_state, const IeTableView* __restrict__ views_all) {
    Smem& sm = *smp;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t t = warp; t < nt; t += NW) {
        const IeTableView tv = (ROUNDS && map_i0 && per_state) ? views_all[(__ldg(map_i0 + t) & IE_AGAIN_INDEX_MASK) / per_state] : tv0;
        const uint32_t* par = &sm.u.seg.out[5 * t];
        const uint32_t mode = par[0], err_g = par[1], olen = par[4];
        if (!olen) continue;
        PieceCopy cp{out + (((uint64_t)par[3] << 32) | par[2]), lane};
        if (mode == 1) cp(tp + sm.t_start[t], olen);
        else if (mode == 2) walk_key_pieces(sm, tv, tp, err_g, cp);
        else if (mode == 3) walk_output_pieces<ROUNDS>(sm, tv, tp, t, cp);
    }
}
__device__ __noinline__ void per_thread_range(Smem* sm, IeTableView tv, const uint8_t* __restrict__ tmpl, const uint64_t* __restrict__ offs, uint64_t i,
                                              uint64_t my_off, bool active, uint64_t r, uint8_t* __restrict__ out, uint64_t out_cap,
                                              uint64_t* __restrict__ out_offs, uint32_t* __restrict__ out_lens, int32_t* __restrict__ status_out,
                                              uint32_t* __restrict__ aux_out, uint32_t* general_list, uint32_t* general_count, uint32_t* overflow,
                                              ie_batch_info* info, uint64_t info_n, uint64_t out_bias) {
    // ---- per-thread exact path for tiles that do not fit the tile tables -----------------------
    uint32_t len = 0, m0 = 0, olen = 0, status = IE_RES_STRING, aux = 0;
    bool verbatim = false;
    const uint8_t* t = tmpl + my_off;
    if (active) {
        const uint64_t b = __ldg(offs + i + 1);
        if (b - my_off > 0x7FFFFFFFull) status = IE_RES_LIMIT;
        else {
            len = (uint32_t)(b - my_off);
            const Prescan ps = prescan(t, len);
            m0 = ps.m0;
            if (ps.punt) status = IE_RES_PUNT;
            else if (ps.n_open == 0) { verbatim = true; olen = len; }
            else fast_traverse<false>(tv, t, len, m0, nullptr, 0, olen, status, aux);
        }
        if (status == IE_RES_PUNT) { olen = 0; general_list[atomicAdd(general_count, 1u)] = (uint32_t)r; }
    }
    uint64_t tile_total16;
    const uint64_t loc = ie_scan::local_scan(sm->scan, olen, 15, &tile_total16);
    const uint64_t off = ie_scan::allocate(sm->scan, &info->out_bytes, tile_total16) + loc;
    if (info_n) info->n = info_n;
    if (!active) return;
    out_offs[r] = off + out_bias; out_lens[r] = olen; status_out[r] = (int32_t)status; aux_out[r] = aux;
    if (olen == 0) return;
    if (off + olen > out_cap) { *overflow = 1u; return; }
    if (verbatim) { uint8_t* wr = out + off; for (uint32_t k = 0; k < len; ++k) wr[k] = __ldg(t + k); }
    else { uint32_t l2, s2, a2; fast_traverse<true>(tv, t, len, m0, out + off + olen, status & 0xFF, l2, s2, a2); }
}

// Resolves the templates [i0, i0 + nt) of one snapshot as one tile.  Returns false (having written nothing) when the
// range outgrows the tile's tables and holds more than IE_SPLIT_MIN templates: the caller retries it in halves.  A range
// of at most IE_SPLIT_MIN templates that still does not fit takes the exact per-thread path instead.
#define IE_SPLIT_MIN 1u
template <bool ROUNDS>
__device__ __forceinline__ bool resolve_range(Smem& sm, const IeTableView& tv_tile, uint32_t state, const uint8_t* __restrict__ tmpl,
                                              const uint64_t* __restrict__ offs, uint64_t n, uint8_t* __restrict__ out, uint64_t out_cap,
                                              uint64_t* __restrict__ out_offs, uint32_t* __restrict__ out_lens,
                                              int32_t* __restrict__ status_out, uint32_t* __restrict__ aux_out, const IeWorkspace& ws,
                                              ie_batch_info* info, uint64_t out_bias, const IeRound& rd, uint32_t tiles_per_state,
                                              uint64_t i0, uint32_t nt) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    PHASE_INIT();
    // The table a template is resolved against: the tile's snapshot, except in a rescan round on a table of several
    // snapshots, where the round's templates come from all of them and each one finds its own through its result index.
    const bool per_template_view = ROUNDS && rd.result_map && rd.per_state;
    auto view_of = [&](uint32_t t) -> IeTableView {
        if (!per_template_view) return tv_tile;
        return rd.views_all[(__ldg(rd.result_map + i0 + t) & IE_AGAIN_INDEX_MASK) / rd.per_state];
    };
    const IeTableView tv = (per_template_view && tid < nt) ? view_of(tid) : tv_tile;  // this thread's own template (P2, P4)
    const uint64_t i = i0 + tid;                    // template
    const bool active = tid < nt;
    // result index; in a rescan round the map also says whether the ORIGINAL template was one whole group (only
    // those may take the typed simple path, interp.rs:45-52 is decided once, on the caller's text)
    const uint32_t mapped = (ROUNDS && rd.result_map && active) ? __ldg(rd.result_map + i) : 0u;
    const uint64_t r = (ROUNDS && rd.result_map) ? (uint64_t)(mapped & IE_AGAIN_INDEX_MASK) : (uint64_t)state * n + i;
    const uint32_t layer_cap = (ROUNDS && rd.result_map) ? mapped >> IE_AGAIN_LAYER_SHIFT : 0xFFFFFFFFu;  // never more layers than the caller's text had
    const bool last_tile = blockIdx.x + 1 == gridDim.x;

    // ---- P0: tile extent ----------------------------------------------------------------------
    const uint64_t off0 = __ldg(offs + i0);
    const uint64_t off_end = __ldg(offs + i0 + nt);
    const uint64_t my_off = active ? __ldg(offs + i) : off_end;
    const uint8_t* __restrict__ tp = tmpl + off0;
    const uint64_t tile_bytes64 = off_end - off0;
    if (tid <= TT) sm.t_start[tid] = (uint32_t)(my_off - off0);
    if (NT == TT && tid == 0) sm.t_start[TT] = (uint32_t)(off_end - off0);
    if (tid < TT) { sm.t_err[tid] = 0; sm.t_flags[tid] = 0; sm.t_splice[tid] = 0; }
    if (tid == 0) { sm.q_n[0] = 0; sm.overflow = 0; sm.ev_n = 0; sm.q_lvl[0] = 0; sm.q_lvl[1] = 0; sm.q_lvl[2] = 0; sm.q_nb = 0; }
    if (tid < 17) {
        auto low = [](int k) -> uint32_t { return k >= 4 ? 0xFFFFFFFFu : k <= 0 ? 0u : (1u << (8 * k)) - 1u; };
        sm.lowmask[tid] = make_uint4(low((int)tid), low((int)tid - 4), low((int)tid - 8), low((int)tid - 12));
    }
    const uintptr_t a0 = (uintptr_t)tp & ~(uintptr_t)15;
    const uint32_t lead = (uint32_t)((uintptr_t)tp - a0);
    const uint32_t tile_bytes = (uint32_t)tile_bytes64;
    const uint32_t n_chunks = (lead + tile_bytes + 15) >> 4;
    const bool too_big = tile_bytes64 + 32 > (uint64_t)M_CAP * 16;  // does not fit the chunk-mask table
    if (too_big && nt > IE_SPLIT_MIN) return false;  // the caller retries with half as many templates

    // ---- P1: flat brace scan --------------------------------------------------------------------
    // Loads are issued P1_BATCH chunks ahead of the compares so that a thread keeps several HBM
    // requests in flight.
#ifdef IE_TMA_TEXT
    // EXPERIMENT (profiles/r02_tma_experiment.md): the tile's text is brought into shared memory by ONE bulk copy of
    // the TMA unit (cp.async.bulk + mbarrier transaction count); the scan then reads its chunks from shared memory.
    extern __shared__ __align__(128) uint8_t tma_text[];
    __shared__ __align__(8) uint64_t tma_bar;
    if (!too_big) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar), dst = (uint32_t)__cvta_generic_to_shared(tma_text);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n_chunks * 16u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(a0),
                         "r"(n_chunks * 16u), "r"(bar)
                         : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
        for (uint32_t c = tid; c < n_chunks; c += NT) {
            const uint4 v = *reinterpret_cast<const uint4*>(tma_text + (size_t)c * 16);
            const uint32_t pv = c ? tma_text[c * 16 - 1] : 0u;  // the byte before the chunk (chunk 0: before the tile, blanked anyway)
            sm.u.scan.cm[c] = scan_chunk(v, c * 16 > lead ? pv : 0u, (int32_t)(c * 16) - (int32_t)lead, tile_bytes, tp);
        }
        __syncthreads();
        if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
    }
#else
    if (!too_big) {
        // (the loop runs per warp, not per lane: the byte before a chunk comes from the neighbouring lane's chunk by
        // shuffle - a byte load per chunk cost as many L1 tag requests as the chunk loads themselves - and only lane 0
        // reads its predecessor from memory)
        for (uint32_t cw = tid & ~31u; cw < n_chunks; cw += NT * P1_BATCH) {
            const uint32_t cb = cw + lane;
            uint4 v[P1_BATCH];
            uint32_t pv[P1_BATCH];
#pragma unroll
            for (int u = 0; u < P1_BATCH; ++u) {
                const uint32_t c = cb + u * NT;
                v[u] = make_uint4(0, 0, 0, 0);
                pv[u] = 0;
                if (c < n_chunks) {
                    v[u] = __ldg(reinterpret_cast<const uint4*>(a0 + (size_t)c * 16));
                    const int32_t p0 = (int32_t)(c * 16) - (int32_t)lead;
                    if (lane == 0 && p0 > 0) pv[u] = __ldg(tp + p0 - 1);  // the byte before the chunk (flat stream)
                }
            }
#pragma unroll
            for (int u = 0; u < P1_BATCH; ++u) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, v[u].w >> 24, 1);
                if (lane) pv[u] = up;
            }
#pragma unroll
            for (int u = 0; u < P1_BATCH; ++u) {
                const uint32_t c = cb + u * NT;
                if (c < n_chunks) sm.u.scan.cm[IE_BOUND(c, M_CAP)] = scan_chunk(v[u], pv[u], (int32_t)(c * 16) - (int32_t)lead, tile_bytes, tp);
            }
        }
    }
#endif
    PHASE_MARK(1);
    __syncthreads();
    PHASE_MARK(2);

    // ---- P2: per-template events and structure ---------------------------------------------------------------
    // One thread per template walks the masks of its own chunks.  A first pass counts its events; a warp scan plus
    // one shared atomic per warp hands every template a contiguous range [eb, eb + ne) of the event arrays (ranges
    // of different templates are in no particular order - nothing downstream relies on one).  The second pass
    // enumerates the events in position order and does the bracket matching on the fly: stack-free through parent
    // links; ev_a[open] counts unresolved children until the group resolves; a group that closes without children
    // is a leaf and goes to the lookup queue right away (P3 skips the queue entries of templates that end up punted).
    if (!too_big && tid < TT) {
        const uint32_t start = sm.t_start[min(tid, nt)], end = sm.t_start[min(tid + 1, nt)];
        const uint32_t ca = lead + start, cz = lead + end;          // the template's extent in chunk coordinates
        const uint32_t c0 = ca >> 4, c1 = active ? (cz + 15) >> 4 : c0;  // its chunks: [c0, c1)
        // valid bytes of the first / last chunk, replicated into both halves of a mask
        const uint32_t keep_first = (0xFFFFu & ~((1u << (ca & 15)) - 1u)) * 0x10001u;
        const uint32_t keep_last = ((2u << ((cz - 1) & 15)) - 1u) * 0x10001u;
        // pass 1: count the template's events; `nonempty` = which of its first 32 chunks hold any
        uint32_t ne = 0, nonempty = 0;
        if (c1 > c0) {
            // first and last chunk with their masks, the chunks between them plain; bit k of `nonempty` = chunk c0 + k
            uint32_t m = sm.u.scan.cm[c0] & keep_first;
            if (c0 + 1 == c1) m &= keep_last;
            uint32_t cnt = __popc((m | (m >> 16)) & 0xFFFFu);
            ne = cnt;
            nonempty = cnt ? 1u : 0u;
            const uint32_t mid_end = min(c1 - 1, c0 + 32);
            uint32_t bit = 2;
            for (uint32_t c = c0 + 1; c < mid_end; ++c, bit <<= 1) {
                const uint32_t mm = sm.u.scan.cm[c];
                ne += __popc((mm | (mm >> 16)) & 0xFFFFu);
                if (mm) nonempty |= bit;
            }
            for (uint32_t c = mid_end; c + 1 < c1; ++c) {  // templates longer than 32 chunks: counted only
                const uint32_t mm = sm.u.scan.cm[c];
                ne += __popc((mm | (mm >> 16)) & 0xFFFFu);
            }
            if (c0 + 1 < c1) {
                m = sm.u.scan.cm[c1 - 1] & keep_last;
                cnt = __popc((m | (m >> 16)) & 0xFFFFu);
                ne += cnt;
                if (cnt && c1 - 1 - c0 < 32) nonempty |= 1u << (c1 - 1 - c0);
            }
        }
        uint32_t incl = ne;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += y;
        }
        uint32_t wbase = 0;
        if (lane == 31 && incl) wbase = atomicAdd(&sm.ev_n, incl);
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
        const uint32_t eb = wbase + incl - ne;
        uint32_t flags = 0, leaf_lo = 0, leaf_top = 0;  // which of the template's first 32 events are leaf groups / leaf groups at the top level
        bool fits = true;
        if (active && eb + ne > (uint32_t)E_CAP) { sm.overflow = 1; fits = false; }  // more brace events than the tile's tables hold
        else if (active) {
            // The flat scan took "the previous byte is a backslash" across template boundaries: a template that
            // starts with a brace right after a template ending in '\\' lost that event -> the general path redoes it.
            if (end > start && start > 0 && __ldg(tp + start - 1) == '\\') {
                const uint8_t b0 = __ldg(tp + start);
                if (b0 == '{' || b0 == '}') flags = TF_PUNT;
            }
            if (flags == 0) {
                // pass 2: ONE loop over the events (every lane runs its `ne` iterations of the same body; a loop over
                // chunks with an inner loop over their events ran at 6 of 32 lanes: templates start at different chunk
                // phases, so at any chunk step only a few lanes have events)
                // The open group and its two nearest ancestors live in registers (par1 = parent of cur_open, par2 = its
                // parent); the links in ev_c[] are written for P3 but read back here only three levels up, off the
                // loop's dependent chain.  Child counts are bumped with fire-and-forget shared atomics.
                uint32_t n_open = 0, cur_open = NONE16, par1 = NONE16, par2 = NONE16;
                bool punt = false, stray = false;
                uint32_t c = c0, m = 0, ev = 0;
                for (uint32_t k = 0; k < ne; ++k) {
                    if (ev == 0) {  // next chunk with events: through the bitmap for the first 32 chunks, by search beyond
                        if (nonempty) { c = c0 + (uint32_t)__ffs(nonempty) - 1u; nonempty &= nonempty - 1u; }
                        else {
                            c = max(c + 1, c0 + 32);
                            for (;; ++c) {
                                uint32_t mm = sm.u.scan.cm[c];
                                if (c + 1 == c1) mm &= keep_last;
                                if (mm) break;
                            }
                        }
                        m = sm.u.scan.cm[c];
                        if (c == c0) m &= keep_first;
                        if (c + 1 == c1) m &= keep_last;
                        ev = (m | (m >> 16)) & 0xFFFFu;
                    }
                    const uint32_t j = (uint32_t)__ffs(ev) - 1u;
                    ev &= ev - 1u;
                    const uint32_t kind = (m >> j) & 0x10001u;  // 1 open, 0x10000 close, both: a byte the tile kernel does not interpret
                    if (kind == 0x10001u) { punt = true; break; }
                    const uint32_t e = eb + k, pos = c * 16 - lead + j;
                    if (kind == 1u) {
                        ++n_open;
                        sm.ev_pos[EI(e)] = pos;
                        sm.ev_c[EI(e)] = (uint16_t)cur_open;
                        sm.ev_a[EI(e)] = 0;
                        if (cur_open != NONE16) atomicAdd(&sm.ev_a[EI(cur_open)], 1u);  // (result unused: no round trip)
                        par2 = par1; par1 = cur_open; cur_open = e;
                    } else {
                        sm.ev_pos[EI(e)] = pos | EV_CLOSE;
                        if (cur_open == NONE16) stray = true;
                        else {
                            const uint32_t o = cur_open;
                            sm.ev_match[EI(o)] = (uint16_t)e; sm.ev_match[EI(e)] = (uint16_t)o;
                            cur_open = par1; par1 = par2;
                            par2 = par1 != NONE16 ? (uint32_t)sm.ev_c[EI(par1)] : NONE16;
                            if (o + 1 == e) {  // closes without children: a leaf group
                                if (k <= 32) { leaf_lo |= 1u << (k - 1); if (cur_open == NONE16) leaf_top |= 1u << (k - 1); }
                                else sm.u.scan.q[IE_BOUND(atomicAdd(&sm.q_n[0], 1u), Q_CAP)] = (tid << 16) | o;
                            }
                        }
                    }
                }
                if (punt) flags = TF_PUNT;
                else if (n_open == 0) flags = TF_VERBATIM;              // the loop at interp.rs:54 is never entered (stray '}' stay)
                else if (stray || cur_open != NONE16) flags = TF_PUNT;  // uneven / improper nesting: general path (exact error text, panic)
                uint32_t layers = 0;
                if (flags == 0 && layer_cap) {
                    // simple-path layers (interp.rs:45-52): leading '{' run matched symmetrically by the trailing '}' run
                    uint32_t ld = 0, tr = 0;
                    while (ld < ne && sm.ev_pos[EI(eb + ld)] == start + ld) ++ld;
                    while (tr < ne && sm.ev_pos[EI(eb + ne - 1 - tr)] == ((end - 1 - tr) | EV_CLOSE)) ++tr;
                    const uint32_t m0 = min(min(ld, tr), layer_cap);
                    for (; layers < m0; ++layers) {
                        if (sm.ev_match[EI(eb + layers)] != eb + ne - 1 - layers) break;
                        sm.ev_pos[EI(eb + layers)] |= EV_SIMPLE;
                    }
                }
                if (ROUNDS) sm.t_layers[tid] = (uint8_t)min(layers, 255u);
            }
            sm.t_eb[tid] = (uint16_t)eb;
            sm.t_ne[tid] = (uint16_t)ne;
            sm.t_flags[tid] = flags;
        }
        // The leaves found above go to the lookup queue, one shared atomic per warp and class instead of one per leaf.
        // Two classes: leaves INSIDE another group enter from the front, leaves at the top level from the back.  P3 hands
        // the queue out in that order, so the threads that will climb a parent chain (`{q-{idx-{slot-A}}}`: three
        // lookups) sit together in the same warps and those warps keep all their lanes busy on every hop; the warps that
        // hold top-level leaves are done after one lookup.  (Mixed, every warp ran its later hops at half its lanes.)
#ifdef IE_P3_LEVELS
        leaf_top = 0;
#endif
        if (!(active && fits)) { leaf_lo = 0; leaf_top = 0; }
        const uint32_t n_in = (uint32_t)__popc(leaf_lo & ~leaf_top), n_top = (uint32_t)__popc(leaf_top);
        uint32_t packed = n_in | (n_top << 16), pincl = packed;  // both counts through one warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, pincl, d);
            if ((int)lane >= d) pincl += y;
        }
        uint32_t base_in = 0, base_top = 0;
        if (lane == 31) {
            if (pincl & 0xFFFFu) base_in = atomicAdd(&sm.q_n[0], pincl & 0xFFFFu);
            if (pincl >> 16) base_top = atomicAdd(&sm.q_nb, pincl >> 16);
        }
        const uint32_t excl = pincl - packed;
        uint32_t qi = __shfl_sync(0xFFFFFFFFu, base_in, 31) + (excl & 0xFFFFu);
        uint32_t qj = __shfl_sync(0xFFFFFFFFu, base_top, 31) + (excl >> 16);
        const uint32_t eb0 = wbase + incl - ne;
        while (leaf_lo) {
            const uint32_t k1 = (uint32_t)__ffs(leaf_lo) - 1u;  // the leaf's close is event k1 + 1 of the template, its open event k1
            leaf_lo &= leaf_lo - 1u;
            const bool top = (leaf_top >> k1) & 1u;
            const uint32_t at = top ? (uint32_t)Q_CAP - 1u - qj : qi;
            sm.u.scan.q[IE_BOUND(at, Q_CAP)] = (tid << 16) | (eb0 + k1);
            if (top) ++qj; else ++qi;
        }
    }
    PHASE_MARK(3);
    __syncthreads();
    PHASE_MARK(4);

    if (sm.overflow && nt > IE_SPLIT_MIN) return false;  // more brace events than the tile's tables hold: half as many templates
    if (too_big || sm.overflow) {
        per_thread_range(&sm, tv, tmpl, offs, i, my_off, active, r, out, out_cap, out_offs, out_lens, status_out, aux_out, ws.general_list,
                         ws.general_count, ws.overflow, info,
                         (tid == 0 && last_tile && !(ROUNDS && rd.n_dev)) ? (uint64_t)gridDim.x / tiles_per_state * n : 0ull, out_bias);
        return true;
    }

    // ---- P3: lookups, one thread per leaf group (and up its parent chain) ------------------------------
    // (IE_P3_LEVELS, experiment: level by level instead.  Level 0 = the leaf groups P2 queued; a thread that resolves the
    // last child of a group appends that group to the queue and the next level starts behind ONE barrier.  The queue is
    // one linear array - a group enters it once - and the levels' sizes rotate through three counters so that every
    // thread sees the same level bounds without a second barrier.)
    {
#ifndef IE_P3_LEVELS
        const uint32_t n_in = sm.q_n[0], nq = n_in + sm.q_nb;  // leaves inside other groups first, then the top-level ones (from the back)
        for (uint32_t k = tid; k < nq; k += NT) {
            const uint32_t item = sm.u.scan.q[k < n_in ? k : (uint32_t)Q_CAP - 1u - (k - n_in)];
            if (sm.t_flags[item >> 16] & (TF_PUNT | TF_VERBATIM)) continue;  // punted after some of its leaves were queued
            if (per_template_view) { const IeTableView tvi = view_of(item >> 16); resolve_group<ROUNDS>(sm, tvi, tp, item >> 16, item & 0xFFFFu); }
            else resolve_group<ROUNDS>(sm, tv_tile, tp, item >> 16, item & 0xFFFFu);
        }
#else
        uint32_t lo = 0, hi = sm.q_n[0];
        for (uint32_t lvl = 0; lo < hi; ++lvl) {
            uint32_t* const next_n = &sm.q_lvl[(lvl + 1) % 3];
            if (tid == 0) sm.q_lvl[(lvl + 2) % 3] = 0;
            for (uint32_t k = lo + tid; k < hi; k += NT) {
                const uint32_t item = sm.u.scan.q[k];
                if (sm.t_flags[item >> 16] & (TF_PUNT | TF_VERBATIM)) continue;  // punted after some of its leaves were queued
                const IeTableView tvi = view_of(item >> 16);
                const uint32_t parent = resolve_group<ROUNDS>(sm, tvi, tp, item >> 16, item & 0xFFFFu);
                if (parent != NONE16) sm.u.scan.q[IE_BOUND(hi + atomicAdd(next_n, 1u), Q_CAP)] = (item & 0xFFFF0000u) | parent;
            }
            __syncthreads();
            lo = hi;
            hi += *next_n;
        }
#endif
    }
    PHASE_MARK(5);
    __syncthreads();
    PHASE_MARK(6);

    // ---- P4: sizes, offsets, copy segments ---------------------------------------------------------------
    uint32_t olen = 0, nseg = 0, status = IE_RES_STRING, aux = 0, err_g = 0;
    uint32_t mode = 0;  // 0 none, 1 verbatim, 2 error key, 3 resolved
    if (active) {
        const uint32_t flags = sm.t_flags[tid];
        const uint32_t err = sm.t_err[tid];
        if (flags & TF_PUNT) {
            status = IE_RES_PUNT;
            ws.general_list[atomicAdd(ws.general_count, 1u)] = (uint32_t)r;
        } else if (flags & TF_VERBATIM) {
            olen = sm.t_start[tid + 1] - sm.t_start[tid];
            nseg = olen ? 1 : 0;
            mode = 1;
        } else if (err && (!ROUNDS || (err >> 8) + 1u > sm.t_splice[tid])) {
            // the rightmost failing group lies right of every spliced value: it is the first failure in the
            // reference's rightmost-first order, and final.  (A failure LEFT of a splice is not: the groups inside
            // the spliced text are resolved first and may fail first - the template goes another round.)
            status = err & 0xFF;
            err_g = err >> 8;
            PieceCount pc;
            walk_key_pieces(sm, tv, tp, err_g, pc);
            olen = pc.bytes; nseg = pc.n;
            mode = 2;
        } else {
            PieceCount pc;
            walk_output_pieces<ROUNDS>(sm, tv, tp, tid, pc);
            olen = pc.bytes; nseg = pc.n;
            mode = 3;
            if (ROUNDS && (flags & TF_AGAIN) && (sm.t_layers[tid] > IE_AGAIN_LAYER_MAX || rd.last_round)) {
                // no further round (or more simple layers than the round map can carry): the general path redoes it
                status = IE_RES_PUNT; olen = 0; nseg = 0; mode = 0;
                ws.general_list[atomicAdd(ws.general_count, 1u)] = (uint32_t)r;
            } else if (ROUNDS && (flags & TF_AGAIN)) {  // its text so far becomes a template of the next round
                status = IE_RES_AGAIN;
                rd.again_list[atomicAdd(rd.again_count, 1u)] = (uint32_t)r | ((uint32_t)sm.t_layers[tid] << IE_AGAIN_LAYER_SHIFT);
                atomicAdd(reinterpret_cast<unsigned long long*>(rd.again_bytes), (unsigned long long)olen);
            } else if (sm.ev_pos[EI(sm.t_eb[tid])] & EV_SIMPLE) {  // the whole template is one group: typed result
                status = IE_RES_TYPED | ((uint32_t)sm.t_tag[tid] << 8);
                aux = sm.t_aux[tid];
            }
        }
    }
    PHASE_MARK(7);
    // Every tile's output starts 16-byte aligned (totals are rounded up), so the chunk structure of the
    // copy sweep does not depend on the tile's global offset: the total is published first, the segment
    // table is built from tile-local offsets, and only then does the look-back collect the predecessors.
    // ONE scan carries both per-template quantities: output bytes in the low 40 bits, copy segments above them.
    constexpr uint64_t LOW40 = (1ull << 40) - 1;
    const uint64_t packed = ((uint64_t)nseg << 40) | olen;
    uint64_t tile_packed;
    const uint64_t excl = ie_scan::local_scan(sm.scan, packed, 0, &tile_packed);
    const uint32_t loc = (uint32_t)(excl & LOW40), sbase = (uint32_t)(excl >> 40), total_seg = (uint32_t)(tile_packed >> 40);
    const uint64_t tile_out64 = tile_packed & LOW40;              // bytes actually produced
    const uint64_t tile_pad64 = (tile_out64 + 15) & ~15ull;       // rounded up to 16: what the tile claims
    const uint32_t tile_out = (uint32_t)tile_out64;
    const uint32_t olead = (uint32_t)((uintptr_t)out & 15);  // tile offsets are multiples of 16
    const uint32_t o_chunks = (olead + tile_out + 15) >> 4;
    const bool seg_ok = total_seg <= (uint32_t)S_CAP && tile_out64 <= 0xFFFFFFFFull;
    const bool index_chunks = o_chunks <= (uint32_t)C_CAP;
    if (seg_ok) {
        if (active && nseg) {
            PieceEmit em{sm, sbase, loc, olead, index_chunks};
            if (mode == 1) em(tp + sm.t_start[tid], olen);
            else if (mode == 2) walk_key_pieces(sm, tv, tp, err_g, em);
            else walk_output_pieces<ROUNDS>(sm, tv, tp, tid, em);
        }
        // chunk 0 starts before the tile's first byte unless the tile's output is 16-byte aligned (then the first piece owns it)
        if (tid == 0) { if (olead) sm.u.seg.cs[0] = (uint16_t)CS_EDGE; sm.u.seg.out[total_seg] = tile_out; }
    }
    PHASE_MARK(8);
    // The tile's output range is claimed with one atomic add on the batch's byte counter: tiles land in
    // the arena in completion order (out_offs[] carries every template's position), so no tile ever
    // waits for a predecessor.
    const uint64_t tile_begin = ie_scan::allocate(sm.scan, &info->out_bytes, tile_pad64);
    const uint64_t tile_end = tile_begin + tile_pad64;
    const uint64_t off = tile_begin + loc;
    if (tid == 0 && last_tile && !(ROUNDS && rd.n_dev)) info->n = (uint64_t)gridDim.x / tiles_per_state * n;
    if (active) {
        out_offs[r] = off + out_bias; out_lens[r] = olen; status_out[r] = (int32_t)status; aux_out[r] = aux;
    }
    if (tile_end > out_cap) { if (tid == 0) *ws.overflow = 1u; return true; }
    if (!seg_ok) {
        // segment table overflow: every thread copies its own pieces
        static_assert(5 * TT <= S_CAP + 2, "the overflow copy keeps five words per template in the segment table");
        __syncthreads();  // (nobody emitted segments: the table is free)
        if (active) {
            uint32_t* par = &sm.u.seg.out[5 * tid];
            par[0] = mode; par[1] = err_g; par[2] = (uint32_t)off; par[3] = (uint32_t)(off >> 32); par[4] = olen;
        }
        __syncthreads();
        copy_own_pieces<ROUNDS>(&sm, tv, tp, nt, out, (ROUNDS && rd.result_map) ? rd.result_map + i0 : nullptr, rd.per_state, rd.views_all);
        return true;
    }
    uint8_t* gout = out + tile_begin;
    const uintptr_t o0 = (uintptr_t)gout & ~(uintptr_t)15;

    PHASE_MARK(9);
    // ---- P5: flat 16-byte output sweep -----------------------------------------------------------------
    // Pass A: every chunk that lies inside ONE segment (constant source misalignment): 5 aligned words,
    // 4 funnel shifts, one 16-byte store.  Pass B: one thread per segment start handles the chunk that
    // contains it (pieces shifted and OR-ed in registers); the ragged first / last chunk of the tile too.
    if (tile_out == 0) return true;
    if (index_chunks) {
        for (uint32_t c = tid; c < o_chunks; c += NT) {
            const uint32_t sidx = sm.u.seg.cs[IE_BOUND(c, C_CAP + 2)];
            if (sidx & CS_EDGE) continue;  // ragged edge of the tile, or a segment ends inside this chunk: pass B
            const uint8_t* src = reinterpret_cast<const uint8_t*>((uintptr_t)sm.u.seg.src[IE_BOUND(sidx, S_CAP)]) + (c * 16 - olead - sm.u.seg.out[sidx]);
            *reinterpret_cast<uint4*>(o0 + (size_t)c * 16) = load16_any(src, 16);
        }
    } else {
        for (uint32_t c = tid; c < o_chunks; c += NT) {
            const int32_t x0s = (int32_t)(c * 16) - (int32_t)olead;  // tile-local output position of the chunk's byte 0
            if (x0s < 0 || (uint32_t)x0s + 16 > tile_out) continue;   // ragged edge chunk: pass B
            const uint32_t xb = (uint32_t)x0s;
            uint32_t lo = 0, hi = total_seg;  // last segment starting at or before xb
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (sm.u.seg.out[mid] <= xb) lo = mid; else hi = mid;
            }
            const uint32_t sidx = lo;
            if (sm.u.seg.out[sidx + 1] < xb + 16) continue;  // a segment starts inside this chunk: pass B
            const uint8_t* src = reinterpret_cast<const uint8_t*>((uintptr_t)sm.u.seg.src[sidx]) + (xb - sm.u.seg.out[sidx]);
            *reinterpret_cast<uint4*>(o0 + (size_t)c * 16) = load16_any(src, 16);
        }
    }
    PHASE_MARK(10);
    // pass B: item 0 = the tile's first chunk, item j >= 1 = the chunk holding the start of segment j when
    // segment j-1 starts at or before that chunk's first byte (the first boundary inside the chunk owns it),
    // item total_seg = the ragged last chunk when no segment start owns it
    for (uint32_t j = tid; j <= total_seg; j += NT) {
        uint32_t c;
        if (j == 0) {
            c = 0;
            if (olead == 0 && sm.u.seg.out[1] >= 16 && tile_out >= 16) continue;  // aligned interior chunk: pass A had it
        } else if (j == total_seg) {
            c = o_chunks - 1;
            const int32_t x0l = (int32_t)(c * 16) - (int32_t)olead;
            if ((uint32_t)(x0l + 16) <= tile_out) continue;                       // last chunk is full: pass A or a boundary item
            if (c == 0 || (int32_t)sm.u.seg.out[j - 1] > x0l) continue;           // item 0 or a boundary item owns it
        } else {
            const uint32_t xo = sm.u.seg.out[j];
            c = (olead + xo) >> 4;
            const int32_t x0j = (int32_t)(c * 16) - (int32_t)olead;
            if (c == 0 || (int32_t)xo == x0j) continue;                            // chunk 0 is item 0's; an aligned start is no boundary
            if ((int32_t)sm.u.seg.out[j - 1] > x0j) continue;                      // an earlier boundary in the same chunk owns it
        }
        const int32_t x0s = (int32_t)(c * 16) - (int32_t)olead;
        const uint32_t xb = x0s < 0 ? 0u : (uint32_t)x0s;
        const uint32_t xe = min(tile_out, (uint32_t)(x0s + 16));
        uint32_t sidx = j ? j - 1 : 0;
        if (j == total_seg) { while (sm.u.seg.out[sidx] > xb) --sidx; }
        uint32_t so = sm.u.seg.out[sidx], se = sm.u.seg.out[sidx + 1];
        // The pieces' bytes land at their place in the chunk through virtual source addresses (load16_range).  The first two
        // pieces (a boundary chunk nearly always has exactly two) are set up together so that their loads are in flight
        // together; a chunk that spans more segments takes the loop.
        uint4 acc = make_uint4(0, 0, 0, 0);
        uint32_t x = xb;
        {
            const uint32_t xn1 = min(xe, se);
            const bool two = xn1 < xe;
            const uint32_t se2 = two ? sm.u.seg.out[sidx + 2] : se;
            const uint32_t xn2 = min(xe, se2);
            const uint8_t* src1 = reinterpret_cast<const uint8_t*>((uintptr_t)sm.u.seg.src[sidx]) + ((int32_t)x0s - (int32_t)so);
            const uint8_t* src2 = reinterpret_cast<const uint8_t*>((uintptr_t)sm.u.seg.src[two ? sidx + 1 : sidx]) + ((int32_t)x0s - (int32_t)se);
            const uint4 v1 = load16_range(sm.lowmask, src1, (uint32_t)((int32_t)x - x0s), (uint32_t)((int32_t)xn1 - x0s));
            uint4 v2 = make_uint4(0, 0, 0, 0);
            if (two) v2 = load16_range(sm.lowmask, src2, (uint32_t)((int32_t)xn1 - x0s), (uint32_t)((int32_t)xn2 - x0s));
            acc.x = v1.x | v2.x; acc.y = v1.y | v2.y; acc.z = v1.z | v2.z; acc.w = v1.w | v2.w;
            x = two ? xn2 : xn1;
            if (two) { ++sidx; so = se; se = se2; }
        }
        while (x < xe) {
            ++sidx; so = se; se = sm.u.seg.out[sidx + 1];
            const uint8_t* src = reinterpret_cast<const uint8_t*>((uintptr_t)sm.u.seg.src[sidx]) + ((int32_t)x0s - (int32_t)so);
            const uint32_t xn = min(xe, se);
            const uint4 v = load16_range(sm.lowmask, src, (uint32_t)((int32_t)x - x0s), (uint32_t)((int32_t)xn - x0s));
            acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
            x = xn;
        }
        if (xe - xb == 16) *reinterpret_cast<uint4*>(o0 + (size_t)c * 16) = acc;
        else {
            for (uint32_t p = xb; p < xe; ++p) {
                const uint32_t q = (uint32_t)((int32_t)p - x0s);
                const uint32_t wq = q < 4 ? acc.x : q < 8 ? acc.y : q < 12 ? acc.z : acc.w;
                gout[p] = (uint8_t)(wq >> (8 * (q & 3)));
            }
        }
    }
    PHASE_MARK(11);
    return true;
}

// Which templates a CTA owns: snapshot, first template and count (a tile = up to tt consecutive templates resolved
// against ONE snapshot, so the table is tile-uniform).  False when the CTA has nothing to do.
template <bool ROUNDS>
__device__ __forceinline__ bool tile_geometry(uint32_t bx, uint32_t& tiles_per_state, uint64_t& n, uint32_t tt, const IeRound& rd,
                                              uint32_t& state, uint64_t& tile_i0, uint32_t& tile_nt) {
    if (ROUNDS && rd.n_dev) {  // a rescan round: the templates are the previous round's unfinished texts, counted on the device
        n = *rd.n_dev;
        if (n == 0) return false;
        // texts grow from round to round (values are spliced in): the tile size follows their mean length
        const uint64_t avg = *rd.bytes_dev / n;
        tt = IE_RESOLVE_TILE;
        while (tt > IE_ROUND_MIN_TILE && (uint64_t)tt * avg * 5 / 4 > IE_TILE_TEXT_BYTES) tt >>= 1;
        if ((uint64_t)bx * tt >= n) return false;
        tiles_per_state = gridDim.x;
    }
    state = bx / tiles_per_state;
    const uint32_t tile = bx - state * tiles_per_state;
    tile_i0 = (uint64_t)tile * tt;
    tile_nt = (uint32_t)min((uint64_t)tt, n - tile_i0);
    return true;
}

// ROUNDS = false is the single-pass kernel (values with groups of their own are punted); ROUNDS = true adds the
// splice bookkeeping of the rescan rounds.  Two instantiations, so the single-pass path pays nothing for it.
template <bool ROUNDS>
__global__ void __launch_bounds__(NT, CTAS_PER_SM) ie_resolve_tile_kernel(const IeTableView* __restrict__ views, uint32_t tiles_per_state,
                                                             const uint8_t* __restrict__ tmpl,
                                                             const uint64_t* __restrict__ offs, uint64_t n, uint8_t* __restrict__ out,
                                                             uint64_t out_cap, uint64_t* __restrict__ out_offs,
                                                             uint32_t* __restrict__ out_lens, int32_t* __restrict__ status_out,
                                                             uint32_t* __restrict__ aux_out, IeWorkspace ws, ie_batch_info* info, uint64_t out_bias,
                                                             uint32_t tt, IeRound rd) {
    __shared__ Smem sm;
    // The tile's templates go through as one range.  A range that outgrows the tile tables (very long or very
    // brace-dense templates) comes back untouched and is retried in halves, down to IE_SPLIT_MIN templates, before the
    // per-thread path takes over.  The retry loop is a second, cold copy of the body which recomputes what it needs
    // (the block index is re-read so that the compiler cannot carry it over), so the common case keeps nothing alive
    // for it.
    uint32_t state, tile_nt;
    uint64_t tile_i0;
    if (!tile_geometry<ROUNDS>(blockIdx.x, tiles_per_state, n, tt, rd, state, tile_i0, tile_nt)) return;
#ifdef IE_L2_PREFETCH
    // One thread asks the TMA unit to pull the text of the tile that will run IE_L2_PREFETCH tiles later (about one
    // wave of resident CTAs) into L2: a single cp.async.bulk.prefetch, no shared memory, no wait.
    if (threadIdx.x == NT - 1 && !(ROUNDS && rd.n_dev)) {
        const uint64_t j0 = tile_i0 + (uint64_t)IE_L2_PREFETCH * tt;
        if (j0 < n) {
            const uint64_t j1 = min(n, j0 + tt);
            const uintptr_t a = ((uintptr_t)tmpl + __ldg(offs + j0)) & ~(uintptr_t)15;
            const uintptr_t b = ((uintptr_t)tmpl + __ldg(offs + j1) + 15) & ~(uintptr_t)15;
            if (b > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((uint32_t)(b - a)) : "memory");
        }
    }
#endif
    {
        const IeTableView tv = views[state];
        if (resolve_range<ROUNDS>(sm, tv, state, tmpl, offs, n, out, out_cap, out_offs, out_lens, status_out, aux_out, ws, info, out_bias, rd,
                                  tiles_per_state, tile_i0, tile_nt))
            return;
    }
    uint32_t bx;
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
    tile_geometry<ROUNDS>(bx, tiles_per_state, n, tt, rd, state, tile_i0, tile_nt);
    const IeTableView tv = views[state];
    uint32_t lo = 0, len = (tile_nt + 1) / 2;
    while (lo < tile_nt) {
        __syncthreads();  // the next range re-initialises the shared tile state
        const uint32_t cur = min(len, tile_nt - lo);
        if (resolve_range<ROUNDS>(sm, tv, state, tmpl, offs, n, out, out_cap, out_offs, out_lens, status_out, aux_out, ws, info, out_bias, rd,
                                  tiles_per_state, tile_i0 + lo, cur))
            lo += cur;
        else
            len = (cur + 1) / 2;
    }
}

}  // namespace

cudaError_t ie_launch_resolve_tiles(const IeTableView* d_views, uint32_t n_states, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                                    uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                                    const IeWorkspace& ws, ie_batch_info* d_info, uint64_t out_bias, uint32_t tt, const IeRound& rd,
                                    cudaStream_t stream) {
    if (rd.n_dev) tt = IE_ROUND_MIN_TILE;       // rounds >= 2: n is the upper bound, the kernel reads the real count and picks tt >= this
    const uint64_t tiles = (n + tt - 1) / tt;
#if defined(IE_TMA_TEXT) || defined(IE_DUMMY_DSMEM)
    // experiment builds: dynamic shared memory for the staged tile text (IE_TMA_TEXT) or the same amount left unused
    // (IE_DUMMY_DSMEM: the occupancy control of the experiment)
    const size_t dsm = (size_t)M_CAP * 16;
    cudaFuncSetAttribute(ie_resolve_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    cudaFuncSetAttribute(ie_resolve_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
#else
    const size_t dsm = 0;
#endif
    if (rd.allow_splice)
        ie_resolve_tile_kernel<true><<<(unsigned)(tiles * n_states), NT, dsm, stream>>>(d_views, (uint32_t)tiles, d_tmpl, d_offs, n, d_out, out_cap,
                                                                                      d_out_offs, d_out_lens, d_status, d_aux, ws, d_info, out_bias,
                                                                                      tt, rd);
    else
    ie_resolve_tile_kernel<false><<<(unsigned)(tiles * n_states), NT, dsm, stream>>>(d_views, (uint32_t)tiles, d_tmpl, d_offs, n, d_out, out_cap, d_out_offs, d_out_lens, d_status,
                                                              d_aux, ws, d_info, out_bias, tt, rd);
    return cudaGetLastError();
}

#ifdef IE_DEBUG_BOUNDS
// out4: violations counted so far, then source line / index / capacity of the first one (this build of the kernel: the
// 128-template and the 32-template build count separately)
#ifdef IE_TILE_SMALL
extern "C" int ie_debug_bound_violations_small(unsigned long long* out4) {
#else
extern "C" int ie_debug_bound_violations(unsigned long long* out4) {
#endif
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out4, g_bound_violations, sizeof(unsigned long long) * 4) == cudaSuccess ? 0 : 1;
}
#endif

#if defined(IE_PHASE_TIMING) && !defined(IE_TILE_SMALL)
extern "C" int ie_debug_phase_cycles(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z); }
    return 0;
}
#endif
