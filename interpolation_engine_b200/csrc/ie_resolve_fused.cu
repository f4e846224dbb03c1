// Fused tile kernel of the batched `{key}` resolver (sm_100a) — the hot path of a launch without rescan rounds.
//
// Replaces interpolate_inserts + get_interpdata (rust-project/src/interp.rs:31-137) for one tile of up to 128 consecutive
// templates per CTA of 128 threads.  The phase-wise kernel (ie_resolve_tile.cu) keeps the bracket structure of a tile in
// shared memory (four event arrays, a lookup queue, child counters) and walks it three times (structure, lookups,
// sizes): 102 of its 208 warp instructions per C4 template, on half of its warps or half of its lanes, behind two
// barriers that cost 18 % of the warp time (profiles/r02s capture).  Here ONE thread takes its template from chunk
// masks to copy segments in a single left-to-right pass over its brace events:
//
//   P1  flat scan      all lanes stream the tile's bytes as 16-byte coalesced chunks; SIMD-in-register byte compares
//                      give one 32-bit brace mask per chunk (ie_tile_common.cuh); two bitmaps by warp ballot: chunks with
//                      any event, chunks with an escaped brace
//   PF  resolve        one thread per template, one loop over its events, the lanes of a warp in lockstep.  The innermost
//                      open group lives in REGISTERS (key so far: 16 bytes, its length, where the next literal piece
//                      starts, where the group opened), the groups around it on an 8-level per-thread stack in local
//                      memory.  '{' appends the pending literal piece to the enclosing key (or stages a literal copy
//                      piece at the top level) and pushes; '}' completes the key, hashes it, probes the table (one
//                      256-bit load = one L2 round trip), applies the type gate (interp.rs:71-80) and pops: the value
//                      goes into the parent's key (inline values arrive with the probe) or becomes a copy piece.
//                      No event arrays, no queue, no shared atomics; a `{q-{idx-{slot-A}}}` chain is three consecutive
//                      iterations of the same thread, and every lane of a warp runs the same body.
//   P4  offsets        one CTA scan (bytes + segment counts), one atomic add claims the tile's arena range, the staged
//                      pieces move into the tile's dense segment table (which takes the chunk masks' place)
//   P5  flat copy      all lanes sweep the tile's output range in 16-byte aligned chunks (as in ie_resolve_tile.cu)
//
// Exactness: the same argument as ie_resolve_tile.cu — with every spliced value free of unescaped braces and sentinel
// corner cases the reference's rightmost-first rewriting equals bracket matching, and the reference's error is the
// failing group with the largest '{' position whose children all succeeded: in close order that is the LAST failing
// group (a group that closes later and is no ancestor opened later; ancestors of a failed group are never looked up).
// The failing key is still in registers when the loop ends and is written out from there.
// What the register pass does not hold is handed on, exactly:
//   * a template that needs more copy pieces than its share of the staging area (S_CAP / templates of the range: 8 in a
//     full tile) counts its pieces and the range comes back to be retried at the size that fits (like a range whose text
//     outgrows the chunk-mask table);
//   * a literal key of more than 16 bytes (no group inside it) is hashed and compared from the tile's text by the byte-wise
//     lookup of ie_device.cuh, in place (its lanes' neighbours wait for it: a long key costs its warp about 1.5 us);
//   * nesting deeper than F_DEPTH levels, a key of more than 16 bytes ASSEMBLED from pieces, a template of more than S_CAP
//     pieces: the exact per-thread traversal of ie_device.cuh at the end of the tile, compacted onto the first lanes;
//   * what the tile kernels never interpret (flagged values, uneven braces, sentinel collisions): the general kernel.
// Launches with rescan rounds (and the 32-template tiles of many snapshots x a handful of templates) stay on
// ie_resolve_tile.cu (ie_resolve.cu picks).
#include <cuda_runtime.h>

#include "ie_common.cuh"
#include "ie_device.cuh"
#include "ie_kernels.h"
#include "ie_scan.cuh"
#include "ie_tile_common.cuh"

namespace {

using namespace ie_dev;
using namespace ie_tile;

// -DIE_DEBUG_BOUNDS: every index into a shared-memory table (and into the per-thread group stack) is checked against the
// table's capacity before it is used; a violation is counted (and the first one recorded) instead of corrupting memory.
// compute-sanitizer is not available on the GPU pool, this build takes its place: tests/fuzz_campaign.py runs against it
// unchanged and ie_debug_bound_violations_fused() must stay 0.
#ifdef IE_DEBUG_BOUNDS
__device__ unsigned long long g_fused_bound_violations[4];  // count, line, index, capacity of the first one
__device__ __noinline__ void fused_bound_report(uint32_t i, uint32_t cap, int line) {
    if (atomicAdd(&g_fused_bound_violations[0], 1ull) == 0) { g_fused_bound_violations[1] = (unsigned long long)line; g_fused_bound_violations[2] = i; g_fused_bound_violations[3] = cap; }
}
__device__ __forceinline__ uint32_t fused_bound_check(uint32_t i, uint32_t cap, int line) {
    if (i >= cap) { fused_bound_report(i, cap, line); return 0u; }
    return i;
}
#define FB(i, cap) fused_bound_check((uint32_t)(i), (uint32_t)(cap), __LINE__)
#else
#define FB(i, cap) (i)
#endif

// cache policy of the table probes (random 64-byte slots of a table far larger than L1: no reuse there)
#ifndef IE_PROBE_HINT
#define IE_PROBE_HINT ".L1::no_allocate"
#endif
#ifndef IE_F_TT
#define IE_F_TT 128
#endif
constexpr int TT = IE_F_TT;  // templates per tile at most (the launch picks tt <= TT from the mean template length)
constexpr int NT = IE_F_TT;  // threads per CTA: one per template
#ifndef IE_F_CTAS
#define IE_F_CTAS (1024 / IE_F_TT)
#endif
#ifndef IE_F_SEGS
#define IE_F_SEGS 8
#endif
#ifndef IE_F_P1_BATCH
#define IE_F_P1_BATCH 4
#endif
constexpr int F_P1_BATCH = IE_F_P1_BATCH;  // chunk loads in flight per thread in P1
constexpr int F_SEGS = IE_F_SEGS;     // copy pieces per template the staging area holds for a full tile
#ifndef IE_F_PA_UNROLL
#define IE_F_PA_UNROLL 2
#endif
constexpr int PA_UNROLL = IE_F_PA_UNROLL;  // output chunks per thread and step of the copy sweep's pass A
constexpr int F_DEPTH = 8;            // nesting depth of the register pass (deeper: per-thread path)
#ifndef IE_F_M_PER
#define IE_F_M_PER IE_M_PER
#endif
constexpr int M_CAP = IE_F_M_PER * TT;  // 16-byte chunks of template text per tile (a longer tile is retried in halves)
constexpr int S_CAP = F_SEGS * TT;    // copy segments per tile: cannot overflow
#ifndef IE_F_C_PER
#define IE_F_C_PER 18
#endif
constexpr int C_CAP = IE_F_C_PER * TT;        // 16-byte output chunks with a segment index (288 bytes of output per template)
constexpr uint32_t CS_EDGE = 0x8000u;  // cs[]: the chunk is not covered by ONE segment (pass B assembles it)
constexpr uint32_t SEG_VALUE = 0x80000000u;  // staged segment (length word): the source is a value of the table (16-byte units from its base)
constexpr uint32_t SEG_TEXT = 0x80000000u;   // segment table (source word): an offset into the tile's text; the launch keeps tables >= 32 GiB off this kernel
constexpr int NZ_WORDS = (M_CAP + IE_F_P1_BATCH * NT) / 32 + 2;  // one bit per chunk a P1 step can touch, + the word a window reads ahead
static_assert(S_CAP < 0x8000, "segment indices share 16 bits with CS_EDGE");

struct SmemF {
    ie_scan::TileSmemT<NT> scan;
    union {
        uint32_t cm[M_CAP];  // P1 -> PF: per chunk, bit j = unescaped '{' at byte j, bit 16 + j = '}', both = punt marker
        struct {
            uint32_t out[S_CAP + 2];  // P4 -> P5: tile-local output offset of each segment (+ sentinel)
            uint32_t src[S_CAP];      //           its source: a value (16-byte units from the table base) or SEG_TEXT | offset in the tile's text
        } seg;
    } u;
    union {
        uint16_t cs[C_CAP + 2];   // P4 -> P5: segment holding the first byte of each 16-byte aligned output chunk | CS_EDGE
        struct {                  // P0 / P1 -> PF (dead once the scan's barrier has passed):
            uint32_t t_start[TT + 1];                            // template start, tile-relative
            uint32_t nz[NZ_WORDS];  // bit c = chunk c holds an event
            uint32_t ez[NZ_WORDS];  // bit c = chunk c holds a brace escaped by the byte before it
        } pf;
    } v;
    uint2 stage[F_SEGS * TT];      // PF -> P4: piece k of template t at [k * TT + t]: (source, length | SEG_VALUE)
    uint4 lowmask[17];             // lowmask[k] = the low k bytes of a 16-byte quantity set (load16_range)
    uint8_t irr[TT];               // templates left to the per-thread path
    uint32_t n_irr;
    uint32_t retry;                // a template outgrew its share of the staging area: the largest range (in templates) that
                                   // gives it enough, 0xFFFFFFFF = none did
};
static_assert(IE_F_CTAS * (sizeof(SmemF) + 1024) <= 196 * 1024, "SmemF outgrew the 196 KB carve-out");

enum : uint32_t { M_SEGS = 0, M_PUNT = 1, M_ERROR = 2, M_IRREGULAR = 3, M_COUNT = 4 };  // (M_COUNT: inside the register pass only)

// One template on the exact per-thread traversal (ie_device.cuh): sizes, claims its own arena range, writes.
__device__ __noinline__ void per_thread_one(IeTableView tv, const uint8_t* __restrict__ tmpl, const uint64_t* __restrict__ offs, uint64_t i, uint64_t r,
                                            uint8_t* __restrict__ out, uint64_t out_cap, uint64_t* __restrict__ out_offs, uint32_t* __restrict__ out_lens,
                                            int32_t* __restrict__ status_out, uint32_t* __restrict__ aux_out, uint32_t* general_list,
                                            uint32_t* general_count, uint32_t* overflow, ie_batch_info* info, uint64_t out_bias) {
    const uint64_t a = __ldg(offs + i), b = __ldg(offs + i + 1);
    const uint8_t* t = tmpl + a;
    uint32_t len = 0, m0 = 0, olen = 0, status = IE_RES_STRING, aux = 0;
    bool verbatim = false;
    if (b - a > 0x7FFFFFFFull) status = IE_RES_LIMIT;
    else {
        len = (uint32_t)(b - a);
        const Prescan ps = prescan(t, len);
        m0 = ps.m0;
        if (ps.punt) status = IE_RES_PUNT;
        else if (ps.n_open == 0) { verbatim = true; olen = len; }
        else fast_traverse<false>(tv, t, len, m0, nullptr, 0, olen, status, aux);
    }
    if (status == IE_RES_PUNT) { olen = 0; general_list[atomicAdd(general_count, 1u)] = (uint32_t)r; }
    uint64_t off = 0;
    if (olen) off = atomicAdd(reinterpret_cast<unsigned long long*>(&info->out_bytes), (unsigned long long)((olen + 15u) & ~15u));
    out_offs[r] = off + out_bias; out_lens[r] = olen; status_out[r] = (int32_t)status; aux_out[r] = aux;
    if (olen == 0) return;
    if (off + olen > out_cap) { *overflow = 1u; return; }
    if (verbatim) { uint8_t* wr = out + off; for (uint32_t k = 0; k < len; ++k) wr[k] = __ldg(t + k); }
    else { uint32_t l2, s2, a2; fast_traverse<true>(tv, t, len, m0, out + off + olen, status & 0xFF, l2, s2, a2); }
}

// A tile that does not fit the chunk-mask table even as a single template: the per-thread path for its templates.
__device__ __noinline__ void per_thread_tile(IeTableView tv, const uint8_t* __restrict__ tmpl, const uint64_t* __restrict__ offs, uint64_t i0, uint32_t nt,
                                             uint64_t r0, uint8_t* __restrict__ out, uint64_t out_cap, uint64_t* __restrict__ out_offs,
                                             uint32_t* __restrict__ out_lens, int32_t* __restrict__ status_out, uint32_t* __restrict__ aux_out,
                                             uint32_t* general_list, uint32_t* general_count, uint32_t* overflow, ie_batch_info* info, uint64_t out_bias) {
    if (threadIdx.x < nt)
        per_thread_one(tv, tmpl, offs, i0 + threadIdx.x, r0 + threadIdx.x, out, out_cap, out_offs, out_lens, status_out, aux_out, general_list, general_count,
                       overflow, info, out_bias);
}

// A literal key longer than 16 bytes (no group inside it): hashed and compared from memory by the byte-wise lookup of
// ie_device.cuh, out of line - the register pass pays for it only where such a key occurs.  Returns the slot index, or
// LONG_MISS / LONG_MISS_ARG (interp.rs:109-116, :136).
constexpr uint32_t LONG_MISS = 0xFFFFFFFFu, LONG_MISS_ARG = 0xFFFFFFFEu;
__device__ __forceinline__ uint32_t lookup_long_key(const uint8_t* base, uint32_t mask, const uint8_t* __restrict__ key, uint32_t len) {
    const IeTableView tv{base, mask, 0u};
    const IeSlot* s = ie_lookup(tv, key, len);
    if (s) return (uint32_t)(s - reinterpret_cast<const IeSlot*>(base));
    return is_arg_key(key, len) ? LONG_MISS_ARG : LONG_MISS;
}

// The literal piece [from, to) of the tile's text appended to a key held in registers.  False: the key outgrows 16 bytes.
__device__ __forceinline__ bool key_append_text(const uint4* __restrict__ lowmask, const uint8_t* __restrict__ tp, uint32_t from, uint32_t to, uint4& key,
                                                uint32_t& klen) {
    const uint32_t m = to - from;
    if (klen + m > 16) return false;
    if (m) {
        const uint4 v = load16_range(lowmask, tp + from - klen, klen, klen + m);
        key.x |= v.x; key.y |= v.y; key.z |= v.z; key.w |= v.w;
        klen += m;
    }
    return true;
}

// Resolves the templates [i0, i0 + nt) of one snapshot as one tile.  Returns false (having written nothing) when the
// range's text outgrows the chunk-mask table, or one of its templates its share of the staging area, and the range holds
// more than one template: the caller retries it in halves.
__device__ __forceinline__ bool resolve_range_fused(SmemF& sm, const IeTableView& tv, uint32_t state, const uint8_t* __restrict__ tmpl,
                                                    const uint64_t* __restrict__ offs, uint64_t n, uint8_t* __restrict__ out, uint64_t out_cap,
                                                    uint64_t* __restrict__ out_offs, uint32_t* __restrict__ out_lens, int32_t* __restrict__ status_out,
                                                    uint32_t* __restrict__ aux_out, const IeWorkspace& ws, ie_batch_info* info, uint64_t out_bias,
                                                    uint32_t tiles_per_state, uint64_t i0, uint32_t nt) {
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint64_t i = i0 + tid;  // this thread's template
    const bool active = tid < nt;
    const uint64_t r = (uint64_t)state * n + i;  // its result index
    const bool last_tile = blockIdx.x + 1 == gridDim.x;

    // ---- P0: tile extent ----------------------------------------------------------------------
    const uint64_t off0 = __ldg(offs + i0);
    const uint64_t off_end = __ldg(offs + i0 + nt);
    const uint64_t my_off = active ? __ldg(offs + i) : off_end;
    const uint8_t* __restrict__ tp = tmpl + off0;
    const uint64_t tile_bytes64 = off_end - off0;
    sm.v.pf.t_start[FB(tid, TT + 1)] = (uint32_t)(my_off - off0);
    if (tid == 0) { sm.v.pf.t_start[FB(TT, TT + 1)] = (uint32_t)(off_end - off0); sm.n_irr = 0; sm.retry = 0xFFFFFFFFu; }
    if (tid < 17) {
        auto low = [](int k) -> uint32_t { return k >= 4 ? 0xFFFFFFFFu : k <= 0 ? 0u : (1u << (8 * k)) - 1u; };
        sm.lowmask[tid] = make_uint4(low((int)tid), low((int)tid - 4), low((int)tid - 8), low((int)tid - 12));
    }
    const uintptr_t a0 = (uintptr_t)tp & ~(uintptr_t)15;
    const uint32_t lead = (uint32_t)((uintptr_t)tp - a0);
    const uint32_t tile_bytes = (uint32_t)tile_bytes64;
    const uint32_t n_chunks = (lead + tile_bytes + 15) >> 4;
    const bool too_big = tile_bytes64 + 32 > (uint64_t)M_CAP * 16;  // does not fit the chunk-mask table
    if (too_big) {
        if (nt > 1) return false;  // the caller retries with half as many templates
        per_thread_tile(tv, tmpl, offs, i0, nt, (uint64_t)state * n + i0, out, out_cap, out_offs, out_lens, status_out, aux_out, ws.general_list,
                        ws.general_count, ws.overflow, info, out_bias);
        if (tid == 0 && last_tile) info->n = (uint64_t)gridDim.x / tiles_per_state * n;
        return true;
    }

    // ---- P1: flat brace scan (ie_resolve_tile.cu P1) --------------------------------------------------------------
    // Besides the mask of every chunk, ONE bit per chunk says whether it holds any event (a warp scans 32 consecutive
    // chunks per step: the ballot of "mask != 0" is that word); PF jumps from event chunk to event chunk through it.  A
    // second bitmap marks the chunks that hold an escaped brace (the first-byte repair of PF looks at text only there).
    for (uint32_t cw = tid & ~31u; cw < n_chunks; cw += NT * F_P1_BATCH) {
        const uint32_t cb = cw + lane;
        uint4 v[F_P1_BATCH];
        uint32_t pv[F_P1_BATCH];
#pragma unroll
        for (int u = 0; u < F_P1_BATCH; ++u) {
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
        for (int u = 0; u < F_P1_BATCH; ++u) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, v[u].w >> 24, 1);
            if (lane) pv[u] = up;
        }
#pragma unroll
        for (int u = 0; u < F_P1_BATCH; ++u) {
            const uint32_t c = cb + u * NT;
            uint32_t mk = 0, esc = 0;
            if (c < n_chunks) {
                mk = scan_chunk(v[u], pv[u], (int32_t)(c * 16) - (int32_t)lead, tile_bytes, tp, &esc);
                sm.u.cm[FB(c, M_CAP)] = mk;
            }
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, mk != 0), eword = __ballot_sync(0xFFFFFFFFu, esc != 0);
            if (lane == 0) { sm.v.pf.nz[FB((cw + u * NT) >> 5, NZ_WORDS)] = word; sm.v.pf.ez[FB((cw + u * NT) >> 5, NZ_WORDS)] = eword; }
        }
    }
    __syncthreads();

    // ---- PF: one thread per template, one pass over its events ------------------------------------------------------
    uint32_t olen = 0, nseg = 0, status = IE_RES_STRING, aux = 0, mode = M_SEGS;
    uint4 EK = make_uint4(0, 0, 0, 0);  // key of the last failing group, its length, its IE_RES_* code
    uint32_t ekl = 0, err_status = 0;
    {
        uint32_t start = 0, end = 0, c0 = 0, c1 = 0, keep_first = 0, keep_last = 0, m0 = 0;
        bool live = false;  // this lane still has events to go through
        if (active) {
            start = sm.v.pf.t_start[FB(tid, TT + 1)]; end = sm.v.pf.t_start[FB(tid + 1, TT + 1)];
            const uint32_t ca = lead + start, cz = lead + end;  // the template's extent in chunk coordinates
            c0 = ca >> 4; c1 = (cz + 15) >> 4;                  // its chunks: [c0, c1)
            // valid bytes of the first / last chunk, replicated into both halves of a mask
            keep_first = (0xFFFFu & ~((1u << (ca & 15)) - 1u)) * 0x10001u;
            keep_last = ((2u << ((cz - 1) & 15)) - 1u) * 0x10001u;
            if (end > start) {
                live = true;
                // The flat scan took "the previous byte is a backslash" across template boundaries: a template that starts
                // with a brace right after a template ending in '\\' lost that event -> the general path redoes it.
                // (only a template whose first chunk holds an escaped brace at all looks at the bytes)
                if (start > 0 && ((sm.v.pf.ez[FB(c0 >> 5, NZ_WORDS)] >> (c0 & 31)) & 1u) && __ldg(tp + start - 1) == '\\') {
                    const uint8_t b0 = __ldg(tp + start);
                    if (b0 == '{' || b0 == '}') { mode = M_PUNT; live = false; }
                }
                // simple-path layers (interp.rs:45-52): the leading '{' run matched symmetrically by the trailing '}' run.
                // Group k of the leading run is a simple layer iff k < min(leading, trailing) and it closes at end - 1 - k
                // (the k bytes behind that close are closes too, so the k groups around it close symmetrically as well).
                auto ev_at = [&](uint32_t p) -> uint32_t { const uint32_t q = lead + p; return (sm.u.cm[FB(q >> 4, M_CAP)] >> (q & 15)) & 0x10001u; };
                if (live && ev_at(start) == 1u && ev_at(end - 1) == 0x10000u) {
                    uint32_t ld = 1, tr = 1;
                    while (start + ld < end && ev_at(start + ld) == 1u) ++ld;
                    while (tr < end - start && ev_at(end - 1 - tr) == 0x10000u) ++tr;
                    m0 = min(ld, tr);
                }
            }
        }
        // The staging area is shared out evenly: S_CAP / nt pieces per template (8 for a full tile).  A template that needs
        // more asks for the range to be retried in halves - 16, 32, ... pieces each - like a range whose text is too long.
        const uint32_t seg_cap = (uint32_t)S_CAP / nt;
        auto stage_piece = [&](uint32_t src, uint32_t len_kind) -> bool {
            if (nseg == seg_cap) return false;
            sm.stage[FB(nseg * nt + tid, F_SEGS * TT)] = make_uint2(src, len_kind);
            ++nseg;
            return true;
        };
        // event chunks [cwin, cwin + 32) of this template, from the per-chunk bitmap of P1
        auto window = [&](uint32_t cwin) -> uint32_t {
            uint32_t bits = __funnelshift_r(sm.v.pf.nz[FB(cwin >> 5, NZ_WORDS)], sm.v.pf.nz[FB((cwin >> 5) + 1, NZ_WORDS)], cwin & 31);
            if (cwin + 32 > c1) bits &= (1u << (c1 - cwin)) - 1u;
            return bits;
        };
        const IeSlot* slots = reinterpret_cast<const IeSlot*>(tv.base);
        // The innermost open group lives in registers: key so far, its length, where its pending literal piece starts,
        // where it opened.  The groups around it wait on a per-thread stack (local memory: one 16-byte store per '{'
        // inside a group, one load per '}').
        uint4 stack_key[F_DEPTH - 1];
        uint32_t stack_meta[F_DEPTH - 1];  // key length | '{' position << 8
        uint4 K0 = make_uint4(0, 0, 0, 0);
        uint32_t kl0 = 0, lit0 = 0, op0 = 0, depth = 0, top_lit = start;
        uint32_t poison = 0;  // bit d: a group inside the open group d levels up failed (that group is never looked up)
        bool seen_open = false, stray = false;
        uint32_t cwin = c0, nzbits = 0, c = 0, m = 0, ev = 0;
        if (live) nzbits = window(c0);
        for (;;) {
            __syncwarp();  // the lanes of a warp walk their events in lockstep: one body per event for all of them
            if (live && ev == 0) {  // next chunk of this template with events
                for (;;) {
                    if (nzbits == 0) {
                        cwin += 32;
                        if (cwin >= c1) { live = false; break; }
                        nzbits = window(cwin);
                        continue;
                    }
                    c = cwin + (uint32_t)__ffs(nzbits) - 1u;
                    nzbits &= nzbits - 1u;
                    m = sm.u.cm[FB(c, M_CAP)];
                    if (c == c0) m &= keep_first;  // (the neighbours' events in a shared chunk)
                    if (c + 1 == c1) m &= keep_last;
                    ev = (m | (m >> 16)) & 0xFFFFu;
                    if (ev) break;
                }
            }
            if (!__any_sync(0xFFFFFFFFu, live)) break;
            if (live) do {
                const uint32_t j = (uint32_t)__ffs(ev) - 1u;
                ev &= ev - 1u;
                const uint32_t kind = (m >> j) & 0x10001u;  // 1 open, 0x10000 close, both: a byte the tile kernels do not interpret
                if (kind == 0x10001u) { mode = M_PUNT; live = false; break; }
                const uint32_t pos = c * 16 - lead + j;
                if (mode == M_COUNT) {
                    // The template has outgrown its share of the staging area: the rest of the pass only counts the pieces
                    // it would stage (in `olen`; structure alone bounds them: no lookups), so that the retry can pick its
                    // range size.
                    if (kind == 1u) { if (depth == 0 && pos > top_lit) ++olen; ++depth; }
                    else if (depth) { if (--depth == 0) { ++olen; top_lit = pos + 1; } }
                    break;
                }
                if (kind == 1u) {
                    seen_open = true;
                    if (depth == 0) {
                        if (pos > top_lit && !stage_piece(top_lit, pos - top_lit)) { mode = M_COUNT; olen = nseg + 1; depth = 1; break; }
                        olen += pos - top_lit;
                    } else {
                        if (!(poison & 1u) && !key_append_text(sm.lowmask, tp, lit0, pos, K0, kl0)) { mode = M_IRREGULAR; live = false; break; }
                        if (depth == (uint32_t)F_DEPTH) { mode = M_IRREGULAR; live = false; break; }
                        stack_key[FB(depth - 1, F_DEPTH - 1)] = K0;
                        stack_meta[FB(depth - 1, F_DEPTH - 1)] = kl0 | (op0 << 8);
                    }
                    K0 = make_uint4(0, 0, 0, 0); kl0 = 0; op0 = pos; lit0 = pos + 1;
                    poison <<= 1;
                    ++depth;
                } else if (depth == 0) stray = true;  // a '}' outside every group: text, unless the template has groups (decided at the end)
                else {
                    bool fail = false;
                    uint32_t vl_tf = 0, val_off16 = 0;
                    uint4 tail_hdr = make_uint4(0, 0, 0, 0), tail_val = make_uint4(0, 0, 0, 0);
                    const uint32_t layer = op0 - start;
                    const bool simple = layer < m0 && pos == end - 1 - layer;
                    if (!(poison & 1u)) {
                        uint32_t err = 0;
                        const bool long_key = kl0 == 0 && pos - lit0 > 16;  // a literal key that does not fit the register quad
                        if (long_key) {
                            const uint32_t si = lookup_long_key(tv.base, tv.mask, tp + lit0, pos - lit0);
                            if (si >= LONG_MISS_ARG) err = si == LONG_MISS_ARG ? IE_RES_ARG_MISSING : IE_RES_NOT_FOUND;
                            else {
                                const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(slots + si));
                                vl_tf = q0.z; val_off16 = q0.w;
                                if (depth > 1 || simple) {
                                    tail_hdr = __ldg(reinterpret_cast<const uint4*>(slots + si) + 2);
                                    tail_val = __ldg(reinterpret_cast<const uint4*>(slots + si) + 3);
                                }
                            }
                        } else if (!key_append_text(sm.lowmask, tp, lit0, pos, K0, kl0)) { mode = M_IRREGULAR; live = false; break; }
                        else if (kl0 == 0) err = IE_RES_EMPTY_KEY;  // interp.rs:105
                        else {
                            const uint32_t h = hash_short(K0, kl0);
                            uint32_t idx = h & tv.mask;
                            const bool want_tail = depth > 1 || simple;  // the parent's key needs the (inline) value, a typed result the entry
                            bool hit = false, probing = true;
                            do {
                                // header and inline key come with ONE 256-bit load: one L2 round trip per probe
                                uint4 q0, q2;
                                asm volatile("ld.global.nc" IE_PROBE_HINT ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                             : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w)
                                             : "l"(slots + idx));
                                if (want_tail)
                                    asm volatile("ld.global.nc" IE_PROBE_HINT ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                                 : "=r"(tail_hdr.x), "=r"(tail_hdr.y), "=r"(tail_hdr.z), "=r"(tail_hdr.w), "=r"(tail_val.x),
                                                   "=r"(tail_val.y), "=r"(tail_val.z), "=r"(tail_val.w)
                                                 : "l"(reinterpret_cast<const uint8_t*>(slots + idx) + 32));
                                hit = q0.x == h && q0.y == kl0 && q2.x == K0.x && q2.y == K0.y && q2.z == K0.z && q2.w == K0.w;
                                probing = !hit && q0.y != IE_SLOT_EMPTY;
                                vl_tf = q0.z; val_off16 = q0.w;
                                idx = (idx + 1) & tv.mask;
                            } while (probing);
                            if (!hit) {  // interp.rs:109-116, :136
                                bool arg = kl0 >= 3 && (K0.x & 0x00FFFFFFu) == 0x00475241u;  // "ARG"
                                for (uint32_t q = 3; arg && q < kl0; ++q) {
                                    const uint32_t wq = q < 4 ? K0.x : q < 8 ? K0.y : q < 12 ? K0.z : K0.w;
                                    const uint32_t b = (wq >> (8 * (q & 3))) & 0xFFu;
                                    arg = b >= '0' && b <= '9';
                                }
                                err = arg ? IE_RES_ARG_MISSING : IE_RES_NOT_FOUND;
                                vl_tf = 0;
                            }
                        }
                        if (!err && !simple) {
                            if (!tag_splices(IE_SLOT_TAG(vl_tf))) err = IE_RES_UNSUPPORTED;  // interp.rs:71-80
                            else if (IE_SLOT_FLAGS(vl_tf) & IE_VF_ANY) { mode = M_PUNT; live = false; break; }  // interp.rs:81-83 would rescan it
                        }
                        // the failing key: the register quad, or (a long literal key) where it stands in the tile's text
                        if (err) { fail = true; err_status = err; if (long_key) { EK.x = lit0; ekl = pos - lit0; } else { EK = K0; ekl = kl0; } }
                    }
                    // pop
                    const bool bad = fail || (poison & 1u);
                    poison >>= 1;
                    --depth;
                    const uint32_t vlen = IE_SLOT_VLEN(vl_tf);
                    if (depth == 0) {
                        if (!bad) {
                            if (vlen && !stage_piece(val_off16, vlen | SEG_VALUE)) { mode = M_COUNT; olen = nseg + 1; top_lit = pos + 1; break; }
                            olen += vlen;
                            if (simple) { status = IE_RES_TYPED | (IE_SLOT_TAG(vl_tf) << 8); aux = tail_hdr.x; }  // the whole template is one group
                        }
                        top_lit = pos + 1;
                    } else {
                        K0 = stack_key[FB(depth - 1, F_DEPTH - 1)];
                        const uint32_t meta = stack_meta[FB(depth - 1, F_DEPTH - 1)];
                        kl0 = meta & 0xFFu; op0 = meta >> 8;
                        if (bad) poison |= 1u;
                        else if (!(poison & 1u)) {
                            // the value joins the parent's key; values of <= 16 bytes came zero-padded with the probe
                            if (kl0 + vlen > 16) { mode = M_IRREGULAR; live = false; break; }
                            or_shifted(K0, tail_val, kl0);
                            kl0 += vlen;
                        }
                        lit0 = pos + 1;
                    }
                }
            } while (0);
        }
        if (mode == M_COUNT) {
            const uint32_t need = olen + (end > top_lit ? 1u : 0u);
            // a range of S_CAP / need templates gives every one of them `need` pieces; more than the whole area: per-thread path
            if (need <= (uint32_t)S_CAP && nt > 1) atomicMin(&sm.retry, max(1u, (uint32_t)S_CAP / need));
            mode = M_IRREGULAR;
        }
        if (active && mode == M_SEGS) {
            if (!seen_open) {  // the loop at interp.rs:54 is never entered (stray '}' stay): verbatim
                olen = end - start;
                if (olen) stage_piece(start, olen);
            } else if (stray || depth != 0) mode = M_PUNT;  // uneven / improper nesting: general path (exact error text, panic)
            else if (err_status) mode = M_ERROR;
            else {
                if (end > top_lit && !stage_piece(top_lit, end - top_lit)) {
                    mode = M_IRREGULAR;
                    if (nseg + 1 <= (uint32_t)S_CAP && nt > 1) atomicMin(&sm.retry, max(1u, (uint32_t)S_CAP / (nseg + 1)));
                }
                olen += end - top_lit;
            }
        }
        if (mode != M_SEGS) { olen = 0; nseg = 0; }
    }
    __syncthreads();
    if (sm.retry < nt) return false;  // (nothing has left the CTA yet; the caller reads the size hint)
    if (mode == M_PUNT) { status = IE_RES_PUNT; aux = 0; ws.general_list[atomicAdd(ws.general_count, 1u)] = (uint32_t)r; }
    else if (mode == M_IRREGULAR) sm.irr[FB(atomicAdd(&sm.n_irr, 1u), TT)] = (uint8_t)tid;

    // ---- P4: offsets and the tile's segment table -----------------------------------------------------------------
    // Every tile's output starts 16-byte aligned (totals are rounded up), so the chunk structure of the copy sweep does
    // not depend on the tile's global offset.  ONE scan carries output bytes (low 40 bits) and copy segments.
    constexpr uint64_t LOW40 = (1ull << 40) - 1;
    const uint64_t packed = ((uint64_t)nseg << 40) | olen;
    uint64_t tile_packed;
    const uint64_t excl = ie_scan::local_scan(sm.scan, packed, 0, &tile_packed);  // (its barrier also ends the chunk masks' life)
    const uint32_t loc = (uint32_t)(excl & LOW40), sbase = (uint32_t)(excl >> 40), total_seg = (uint32_t)(tile_packed >> 40);
    const uint64_t tile_out64 = tile_packed & LOW40;         // bytes actually produced
    const uint64_t tile_pad64 = (tile_out64 + 15) & ~15ull;  // rounded up to 16: what the tile claims
    const uint32_t olead = (uint32_t)((uintptr_t)out & 15);  // tile offsets are multiples of 16
    // a tile of more than 4 GiB of output (values of up to 32 MiB each) leaves every template to the per-thread path
    const bool huge = tile_out64 > 0xFFFFFFFFull;
    const uint32_t tile_out = huge ? 0u : (uint32_t)tile_out64;
    const uint32_t o_chunks = (olead + tile_out + 15) >> 4;
    const bool index_chunks = o_chunks <= (uint32_t)C_CAP;
    if (!huge && nseg) {
        uint32_t off = loc;
        for (uint32_t k = 0; k < nseg; ++k) {
            const uint2 pc = sm.stage[FB(k * nt + tid, F_SEGS * TT)];
            const uint32_t len = pc.y & ~SEG_VALUE, idx = sbase + k;
            sm.u.seg.out[FB(idx, S_CAP + 2)] = off;
            sm.u.seg.src[FB(idx, S_CAP)] = (pc.y & SEG_VALUE) ? pc.x : (pc.x | SEG_TEXT);
            if (index_chunks) {
                // every 16-byte aligned output chunk whose first byte lies in this piece points back at it; only the last
                // of them can reach beyond the piece's end (CS_EDGE: pass B assembles that chunk)
                const uint32_t lo = off + olead, hi = lo + len;  // the piece in chunk coordinates
                // (pieces of C4-like batches span up to seven chunks: eight predicated stores instead of a loop whose trip
                // count differs from lane to lane; longer pieces finish in the loop)
                const uint32_t cf = (lo + 15) >> 4, ce = hi >> 4;  // chunks [cf, ce) lie inside the piece
                uint16_t* row = &sm.v.cs[0] + cf;
#pragma unroll
                for (int q = 0; q < 8; ++q) if (cf + q < ce) row[FB(cf + q, C_CAP + 2) - cf] = (uint16_t)idx;
                for (uint32_t c = cf + 8; c < ce; ++c) sm.v.cs[FB(c, C_CAP + 2)] = (uint16_t)idx;
                const uint32_t cl = max(cf, ce);
                if ((cl << 4) < hi) sm.v.cs[FB(cl, C_CAP + 2)] = (uint16_t)(idx | CS_EDGE);
            }
            off += len;
        }
    }
    // chunk 0 starts before the tile's first byte unless the tile's output is 16-byte aligned (then the first piece owns it)
    if (tid == 0) { if (olead) sm.v.cs[FB(0, C_CAP + 2)] = (uint16_t)CS_EDGE; sm.u.seg.out[FB(total_seg, S_CAP + 2)] = tile_out; }
    // The tile's output range is claimed with one atomic add on the batch's byte counter: tiles land in the arena in
    // completion order (out_offs[] carries every template's position), so no tile ever waits for a predecessor.
    const uint64_t tile_begin = ie_scan::allocate(sm.scan, &info->out_bytes, huge ? 0ull : tile_pad64);  // (its barrier publishes the segment table)
    const uint64_t tile_end = tile_begin + (huge ? 0ull : tile_pad64);
    if (tid == 0 && last_tile) info->n = (uint64_t)gridDim.x / tiles_per_state * n;
    if (active && !huge && mode <= M_PUNT) {
        out_offs[r] = tile_begin + loc + out_bias; out_lens[r] = olen; status_out[r] = (int32_t)status; aux_out[r] = aux;
    }
    if (active && mode == M_ERROR) {
        // the failing key goes out from the registers it was assembled in (a long literal key: from the tile's text), into a
        // range of its own
        uint64_t off = 0;
        const uint32_t claim = (ekl + 15u) & ~15u;
        if (ekl) off = atomicAdd(reinterpret_cast<unsigned long long*>(&info->out_bytes), (unsigned long long)claim);
        out_offs[r] = off + out_bias; out_lens[r] = ekl; status_out[r] = (int32_t)err_status; aux_out[r] = 0;
        if (ekl) {
            if (off + claim > out_cap) *ws.overflow = 1u;
            else if (ekl > 16) { for (uint32_t q = 0; q < ekl; ++q) out[off + q] = __ldg(tp + EK.x + q); }
            else if (((uintptr_t)(out + off) & 15) == 0) *reinterpret_cast<uint4*>(out + off) = EK;
            else for (uint32_t q = 0; q < ekl; ++q) {
                const uint32_t wq = q < 4 ? EK.x : q < 8 ? EK.y : q < 12 ? EK.z : EK.w;
                out[off + q] = (uint8_t)(wq >> (8 * (q & 3)));
            }
        }
    }
    const uint32_t n_irr = sm.n_irr;
    if (tile_end > out_cap) { if (tid == 0) *ws.overflow = 1u; }
    else if (tile_out) {
        // ---- P5: flat 16-byte output sweep (ie_resolve_tile.cu P5) ------------------------------------------------
        auto seg_src = [&](uint32_t idx) -> uintptr_t {
            const uint32_t raw = sm.u.seg.src[FB(idx, S_CAP)];
#ifdef IE_F_DECODE2
            const bool txt = (int32_t)raw < 0;  // one multiply-add on a selected base and scale
            return (txt ? (uintptr_t)tp : (uintptr_t)tv.base) + (uint64_t)(raw & ~SEG_TEXT) * (txt ? 1u : 16u);
#else
            return (raw & SEG_TEXT) ? (uintptr_t)tp + (raw & ~SEG_TEXT) : (uintptr_t)tv.base + (uintptr_t)raw * 16u;
#endif
        };
        uint8_t* gout = out + tile_begin;
        const uintptr_t o0 = (uintptr_t)gout & ~(uintptr_t)15;
        // Pass A: every chunk that lies inside ONE segment: two aligned loads, register selects, one 16-byte store.
        if (index_chunks) {
            // PA_UNROLL chunks per thread and step: their segment lookups first, then all their loads, then the stores.  The
            // lanes of a warp hold consecutive chunks; where the next lane reads on in the same source (its address is mine
            // + 16), ITS first block is my second one: it comes by shuffle instead of a second load (the L1 data pipe, not
            // the issue slots, bounds this kernel; about two in three second loads go away).
            for (uint32_t cw = tid & ~31u; cw < o_chunks; cw += NT * PA_UNROLL) {
                uintptr_t sa[PA_UNROLL];
                bool ok[PA_UNROLL];
#pragma unroll
                for (int u = 0; u < PA_UNROLL; ++u) {
                    const uint32_t c = cw + lane + u * NT;
                    const uint32_t sidx = c < o_chunks ? sm.v.cs[FB(c, C_CAP + 2)] : CS_EDGE;
                    ok[u] = !(sidx & CS_EDGE);  // (edge: ragged edge of the tile, or a segment ends inside this chunk: pass B)
                    sa[u] = 0;
                    if (ok[u]) sa[u] = seg_src(sidx) + (c * 16 - olead - sm.u.seg.out[FB(sidx, S_CAP + 2)]);
                }
                uint4 A[PA_UNROLL], B[PA_UNROLL];
#pragma unroll
                for (int u = 0; u < PA_UNROLL; ++u) {
                    A[u] = make_uint4(0, 0, 0, 0);
                    if (ok[u]) A[u] = __ldg(reinterpret_cast<const uint4*>(sa[u] & ~(uintptr_t)15));
                }
#pragma unroll
                for (int u = 0; u < PA_UNROLL; ++u) {
#ifdef IE_F_PA_SHFL
                    const uint32_t lo_next = __shfl_down_sync(0xFFFFFFFFu, (uint32_t)sa[u], 1), hi_next = __shfl_down_sync(0xFFFFFFFFu, (uint32_t)(sa[u] >> 32), 1);
                    B[u].x = __shfl_down_sync(0xFFFFFFFFu, A[u].x, 1); B[u].y = __shfl_down_sync(0xFFFFFFFFu, A[u].y, 1);
                    B[u].z = __shfl_down_sync(0xFFFFFFFFu, A[u].z, 1); B[u].w = __shfl_down_sync(0xFFFFFFFFu, A[u].w, 1);
                    const bool from_next = lane < 31 && (((uint64_t)hi_next << 32) | lo_next) == (uint64_t)sa[u] + 16;
                    if (ok[u] && (sa[u] & 15) && !from_next) B[u] = __ldg(reinterpret_cast<const uint4*>(sa[u] & ~(uintptr_t)15) + 1);
#else
                    B[u] = make_uint4(0, 0, 0, 0);
                    if (ok[u] && (sa[u] & 15)) B[u] = __ldg(reinterpret_cast<const uint4*>(sa[u] & ~(uintptr_t)15) + 1);
#endif
                }
#pragma unroll
                for (int u = 0; u < PA_UNROLL; ++u) {
                    if (ok[u]) *reinterpret_cast<uint4*>(o0 + (size_t)(cw + lane + u * NT) * 16) = align16(A[u], B[u], (uint32_t)(sa[u] & 15));
                }
            }
        } else {
            for (uint32_t c = tid; c < o_chunks; c += NT) {
                const int32_t x0s = (int32_t)(c * 16) - (int32_t)olead;  // tile-local output position of the chunk's byte 0
                if (x0s < 0 || (uint32_t)x0s + 16 > tile_out) continue;   // ragged edge chunk: pass B
                const uint32_t xb = (uint32_t)x0s;
                uint32_t lo = 0, hi = total_seg;  // last segment starting at or before xb
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (sm.u.seg.out[FB(mid, S_CAP + 2)] <= xb) lo = mid; else hi = mid;
                }
                const uint32_t sidx = lo;
                if (sm.u.seg.out[FB(sidx + 1, S_CAP + 2)] < xb + 16) continue;  // a segment starts inside this chunk: pass B
                const uint8_t* src = reinterpret_cast<const uint8_t*>(seg_src(sidx)) + (xb - sm.u.seg.out[FB(sidx, S_CAP + 2)]);
                *reinterpret_cast<uint4*>(o0 + (size_t)c * 16) = load16_any(src, 16);
            }
        }
        // Pass B: item 0 = the tile's first chunk, item j >= 1 = the chunk holding the start of segment j when segment
        // j-1 starts at or before that chunk's first byte (the first boundary inside the chunk owns it), item total_seg =
        // the ragged last chunk when no segment start owns it
        for (uint32_t j = tid; j <= total_seg; j += NT) {
            uint32_t c;
            if (j == 0) {
                c = 0;
                if (olead == 0 && sm.u.seg.out[FB(1, S_CAP + 2)] >= 16 && tile_out >= 16) continue;  // aligned interior chunk: pass A had it
            } else if (j == total_seg) {
                c = o_chunks - 1;
                const int32_t x0l = (int32_t)(c * 16) - (int32_t)olead;
                if ((uint32_t)(x0l + 16) <= tile_out) continue;                       // last chunk is full: pass A or a boundary item
                if (c == 0 || (int32_t)sm.u.seg.out[FB(j - 1, S_CAP + 2)] > x0l) continue;           // item 0 or a boundary item owns it
            } else {
                const uint32_t xo = sm.u.seg.out[FB(j, S_CAP + 2)];
                c = (olead + xo) >> 4;
                const int32_t x0j = (int32_t)(c * 16) - (int32_t)olead;
                if (c == 0 || (int32_t)xo == x0j) continue;                            // chunk 0 is item 0's; an aligned start is no boundary
                if ((int32_t)sm.u.seg.out[FB(j - 1, S_CAP + 2)] > x0j) continue;                      // an earlier boundary in the same chunk owns it
            }
            const int32_t x0s = (int32_t)(c * 16) - (int32_t)olead;
            const uint32_t xb = x0s < 0 ? 0u : (uint32_t)x0s;
            const uint32_t xe = min(tile_out, (uint32_t)(x0s + 16));
            uint32_t sidx = j ? j - 1 : 0;
            if (j == total_seg) { while (sm.u.seg.out[FB(sidx, S_CAP + 2)] > xb) --sidx; }
            uint32_t so = sm.u.seg.out[FB(sidx, S_CAP + 2)], se = sm.u.seg.out[FB(sidx + 1, S_CAP + 2)];
            // The pieces' bytes land at their place in the chunk through virtual source addresses (load16_range).  The
            // first two pieces (a boundary chunk nearly always has exactly two) are set up together so that their loads
            // are in flight together; a chunk that spans more segments takes the loop.
            uint4 acc = make_uint4(0, 0, 0, 0);
            uint32_t x = xb;
            {
                const uint32_t xn1 = min(xe, se);
                const bool two = xn1 < xe;
                const uint32_t se2 = two ? sm.u.seg.out[FB(sidx + 2, S_CAP + 2)] : se;
                const uint32_t xn2 = min(xe, se2);
                const uint8_t* src1 = reinterpret_cast<const uint8_t*>(seg_src(sidx)) + ((int32_t)x0s - (int32_t)so);
                const uint8_t* src2 = reinterpret_cast<const uint8_t*>(seg_src(two ? sidx + 1 : sidx)) + ((int32_t)x0s - (int32_t)se);
                const uint4 v1 = load16_range(sm.lowmask, src1, (uint32_t)((int32_t)x - x0s), (uint32_t)((int32_t)xn1 - x0s));
                uint4 v2 = make_uint4(0, 0, 0, 0);
                if (two) v2 = load16_range(sm.lowmask, src2, (uint32_t)((int32_t)xn1 - x0s), (uint32_t)((int32_t)xn2 - x0s));
                acc.x = v1.x | v2.x; acc.y = v1.y | v2.y; acc.z = v1.z | v2.z; acc.w = v1.w | v2.w;
                x = two ? xn2 : xn1;
                if (two) { ++sidx; so = se; se = se2; }
            }
            while (x < xe) {
                ++sidx; so = se; se = sm.u.seg.out[FB(sidx + 1, S_CAP + 2)];
                const uint8_t* src = reinterpret_cast<const uint8_t*>(seg_src(sidx)) + ((int32_t)x0s - (int32_t)so);
                const uint32_t xn = min(xe, se);
                const uint4 v = load16_range(sm.lowmask, src, (uint32_t)((int32_t)x - x0s), (uint32_t)((int32_t)xn - x0s));
                acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
                x = xn;
            }
            // (a 16-byte aligned arena: the ragged tail of the tile's last chunk lies in the padding the tile claimed itself)
            if (xe - xb == 16 || olead == 0) *reinterpret_cast<uint4*>(o0 + (size_t)c * 16) = acc;
            else {
                for (uint32_t p = xb; p < xe; ++p) {
                    const uint32_t q = (uint32_t)((int32_t)p - x0s);
                    const uint32_t wq = q < 4 ? acc.x : q < 8 ? acc.y : q < 12 ? acc.z : acc.w;
                    gout[p] = (uint8_t)(wq >> (8 * (q & 3)));
                }
            }
        }
    }
    // ---- what the register pass left: the exact per-thread traversal, compacted onto the first lanes ---------------
    if (huge ? (active && (mode == M_SEGS || mode == M_IRREGULAR)) : tid < n_irr) {
        const uint32_t t = huge ? tid : sm.irr[FB(tid, TT)];
        per_thread_one(tv, tmpl, offs, i0 + t, (uint64_t)state * n + i0 + t, out, out_cap, out_offs, out_lens, status_out, aux_out, ws.general_list,
                       ws.general_count, ws.overflow, info, out_bias);
    }
    return true;
}

__global__ void __launch_bounds__(NT, IE_F_CTAS) ie_resolve_fused_kernel(const IeTableView* __restrict__ views, uint32_t tiles_per_state,
                                                                        const uint8_t* __restrict__ tmpl, const uint64_t* __restrict__ offs, uint64_t n,
                                                                        uint8_t* __restrict__ out, uint64_t out_cap, uint64_t* __restrict__ out_offs,
                                                                        uint32_t* __restrict__ out_lens, int32_t* __restrict__ status_out,
                                                                        uint32_t* __restrict__ aux_out, IeWorkspace ws, ie_batch_info* info, uint64_t out_bias,
                                                                        uint32_t tt) {
    __shared__ SmemF sm;
    // A tile = up to tt consecutive templates resolved against ONE snapshot.  A range whose text outgrows the chunk-mask
    // table comes back untouched and is retried in halves (a second, cold copy of the body), down to single templates.
    const uint32_t state = blockIdx.x / tiles_per_state;
    const uint32_t tile = blockIdx.x - state * tiles_per_state;
    const uint64_t tile_i0 = (uint64_t)tile * tt;
    const uint32_t tile_nt = (uint32_t)min((uint64_t)tt, n - tile_i0);
    const IeTableView tv = views[state];
    if (resolve_range_fused(sm, tv, state, tmpl, offs, n, out, out_cap, out_offs, out_lens, status_out, aux_out, ws, info, out_bias, tiles_per_state, tile_i0,
                            tile_nt))
        return;
    uint32_t lo = 0, len = tile_nt;
    while (lo < tile_nt) {
        // a range that came back: half as many templates, or what its densest template asked for; after a range that went
        // through the size doubles again (the dense template is behind us)
        __syncthreads();
        const uint32_t hint = sm.retry;
        __syncthreads();  // the next range re-initialises the shared tile state
        len = min((len + 1) / 2, hint);
        for (;;) {
            const uint32_t cur = min(len, tile_nt - lo);
            if (!resolve_range_fused(sm, tv, state, tmpl, offs, n, out, out_cap, out_offs, out_lens, status_out, aux_out, ws, info, out_bias,
                                     tiles_per_state, tile_i0 + lo, cur))
                break;
            lo += cur;
            if (lo >= tile_nt) break;
            len = min(len * 2, tile_nt);
            __syncthreads();
        }
    }
}

}  // namespace

cudaError_t ie_launch_resolve_fused(const IeTableView* d_views, uint32_t n_states, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                                    uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                                    const IeWorkspace& ws, ie_batch_info* d_info, uint64_t out_bias, uint32_t tt, cudaStream_t stream) {
    if (TT > IE_RESOLVE_TILE && tt == IE_RESOLVE_TILE) tt = TT;  // (experiment builds with larger / smaller tiles)
    if (tt > (uint32_t)TT) tt = TT;
    const uint64_t tiles = (n + tt - 1) / tt;
    ie_resolve_fused_kernel<<<(unsigned)(tiles * n_states), NT, 0, stream>>>(d_views, (uint32_t)tiles, d_tmpl, d_offs, n, d_out, out_cap, d_out_offs,
                                                                          d_out_lens, d_status, d_aux, ws, d_info, out_bias, tt);
    return cudaGetLastError();
}

#ifdef IE_DEBUG_BOUNDS
// out4: violations counted so far, then source line / index / capacity of the first one
extern "C" int ie_debug_bound_violations_fused(unsigned long long* out4) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out4, g_fused_bound_violations, sizeof(unsigned long long) * 4) == cudaSuccess ? 0 : 1;
}
#endif
