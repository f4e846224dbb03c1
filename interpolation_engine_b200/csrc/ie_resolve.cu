// Batched `{key}` resolver kernels for sm_100a.
//
// Replaces interpolate_inserts (rust-project/src/interp.rs:31-89) + get_interpdata (:91-137) for
// many independent templates against one immutable inserts snapshot.
//
//   ie_resolve_tile_kernel     (ie_resolve_tile.cu) the hot path: cooperative CTA tiles, single pass
//                              over HBM, compacted output.
//   ie_resolve_general_kernel  exact right-to-left rewriting machine for the templates the tile
//                              kernel declines (values that are rescanned, sentinel collisions,
//                              uneven braces, very long keys, deep nesting).
//   ie_lookup_kernel           get_interpdata for literal keys.
#include <cuda_runtime.h>

#include "ie_common.cuh"
#include "ie_kernels.h"
#include "ie_device.cuh"
#include "ie_scan.cuh"

namespace {

using namespace ie_dev;
constexpr uint32_t MAXF = 24;   // splice depth of the general path's first tier (frames in registers / local memory); deeper
                                // stacks go to the full-size tier, whose frames live in its global scratch

// ---- general path ----------------------------------------------------------------------------
struct Frame {
    const uint8_t* ptr;
    uint32_t len, pos;
};

__device__ void count_braces(const uint8_t* v, uint32_t len, uint32_t& opens, uint32_t& closes) {
    uint8_t prev = 0;
    for (uint32_t i = 0; i < len; ++i) {
        const uint8_t c = v[i];
        if (prev != '\\') { if (c == '{') ++opens; else if (c == '}') ++closes; }
        prev = c;
    }
}

// get_simple_insertkey (interp.rs:11-29) on the sentinelised form of t[lo, hi): escaped braces are
// sentinel text there, i.e. ordinary middle characters.
__device__ bool is_simple_range(const uint8_t* t, uint32_t lo, uint32_t hi) {
    if (hi - lo < 2 || t[lo] != '{' || t[hi - 1] != '}') return false;
    if (hi - lo >= 3 && t[hi - 2] == '\\') return false;  // last char is part of "〠."
    int depth = 0;
    uint8_t prev = 0;
    for (uint32_t i = lo; i < hi; ++i) {
        const uint8_t c = t[i];
        const bool esc = (i > lo) && prev == '\\';
        if (c == '}' && !esc) --depth;
        if ((depth == 0) != (i == lo || i == hi - 1)) return false;
        if (c == '{' && !esc) ++depth;
        prev = c;
    }
    return true;
}

// unsentinelise (interp.rs:86-87: replace ".〠" -> "\{" everywhere, then "〠." -> "\}"), streamed.
// An "〠." survives the first pass iff its '.' does not start a ".〠" (see DESIGN.md).
template <bool WRITE>
__device__ uint32_t unsentinelise(const uint8_t* b, uint32_t len, uint8_t* dst) {
    uint32_t o = 0, i = 0;
    while (i < len) {
        if (i + 4 <= len) {
            const uint8_t b0 = b[i], b1 = b[i + 1], b2 = b[i + 2], b3 = b[i + 3];
            if (b0 == 0x2E && b1 == 0xE3 && b2 == 0x80 && b3 == 0xA0) {
                if (WRITE) { dst[o] = '\\'; dst[o + 1] = '{'; }
                o += 2; i += 4;
                continue;
            }
            if (b0 == 0xE3 && b1 == 0x80 && b2 == 0xA0 && b3 == 0x2E &&
                !(i + 7 <= len && b[i + 4] == 0xE3 && b[i + 5] == 0x80 && b[i + 6] == 0xA0)) {
                if (WRITE) { dst[o] = '\\'; dst[o + 1] = '}'; }
                o += 2; i += 4;
                continue;
            }
        }
        if (WRITE) dst[o] = b[i];
        ++o; ++i;
    }
    return o;
}

// sentinelise (interp.rs:42-43) of raw bytes v[0, len), streamed; the first byte is never escaped
// by what precedes it (see DESIGN.md "general path").
template <bool WRITE>
__device__ uint32_t sentinelise(const uint8_t* v, uint32_t len, uint8_t* dst) {
    uint32_t o = 0;
    for (uint32_t i = 0; i < len; ++i) {
        const uint8_t c = v[i];
        if (c == '\\' && i + 1 < len && (v[i + 1] == '{' || v[i + 1] == '}')) {
            if (WRITE) {
                if (v[i + 1] == '{') { dst[o] = 0x2E; dst[o + 1] = 0xE3; dst[o + 2] = 0x80; dst[o + 3] = 0xA0; }
                else { dst[o] = 0xE3; dst[o + 1] = 0x80; dst[o + 2] = 0xA0; dst[o + 3] = 0x2E; }
            }
            o += 4; ++i;
            continue;
        }
        if (WRITE) dst[o] = c;
        ++o;
    }
    return o;
}

// ---- warp-wide bulk passes of the general path ---------------------------------------------------------------------
// The machine itself is serial (lane 0), but what surrounds it is not: counting the braces of the template and writing
// the result (a copy, or the sentinelised text of the "uneven" error) are per-byte decisions that only look at the
// neighbouring byte.  All 32 lanes do those (measured on 1 KiB texts: 95 % of the kernel's instructions were in them).
constexpr uint32_t FULL = 0xFFFFFFFFu;

__device__ __forceinline__ void warp_count_braces(const uint8_t* v, uint32_t len, uint32_t lane, uint32_t& opens, uint32_t& closes) {
    uint32_t o = 0, c = 0;
    for (uint32_t i = lane; i < len; i += 32) {
        const uint8_t b = v[i];
        if ((b == '{' || b == '}') && !(i > 0 && v[i - 1] == '\\')) { if (b == '{') ++o; else ++c; }
    }
    opens += __reduce_add_sync(FULL, o);
    closes += __reduce_add_sync(FULL, c);
}

// sentinelise() above, 32 bytes per step.  "\{" / "\}" pairs cannot overlap (the second byte is a brace, the first a
// backslash), so whether byte i starts a pair, ends one or stands alone is decided by its neighbours; a warp scan of
// the emitted sizes places the output.  dst == nullptr: length only.
__device__ __forceinline__ uint32_t warp_sentinelise(const uint8_t* v, uint32_t len, uint8_t* dst, uint32_t lane) {
    uint32_t total = 0;
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t i = base + lane;
        uint8_t c = 0, nx = 0;
        bool second = false;
        if (i < len) {
            c = v[i];
            if (c == '\\' && i + 1 < len) nx = v[i + 1];
            second = (c == '{' || c == '}') && i > 0 && v[i - 1] == '\\';
        }
        const bool first = nx == '{' || nx == '}';
        const uint32_t emit = i < len ? (first ? 4u : second ? 0u : 1u) : 0u;
        uint32_t incl = emit;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, d); if (lane >= (uint32_t)d) incl += y; }
        if (dst && emit) {
            uint8_t* w = dst + total + (incl - emit);
            if (first) {
                if (nx == '{') { w[0] = 0x2E; w[1] = 0xE3; w[2] = 0x80; w[3] = 0xA0; }
                else { w[0] = 0xE3; w[1] = 0x80; w[2] = 0xA0; w[3] = 0x2E; }
            } else w[0] = c;
        }
        total += __shfl_sync(FULL, incl, 31);
    }
    return total;
}

__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane) {
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}
__device__ __forceinline__ bool warp_has_byte(const uint8_t* src, uint32_t n, uint8_t what, uint32_t lane) {
    bool hit = false;
    for (uint32_t i = lane; i < n; i += 32) hit |= src[i] == what;
    return __any_sync(FULL, hit);
}
template <typename P>
__device__ __forceinline__ P* warp_bcast_ptr(P* p) {
    return reinterpret_cast<P*>(__shfl_sync(FULL, (unsigned long long)reinterpret_cast<uintptr_t>(p), 0));
}

__device__ uint8_t* reserve_out(uint8_t* out, uint64_t out_cap, ie_batch_info* info, uint32_t* overflow, uint32_t len,
                                uint64_t& off) {
    off = atomicAdd(reinterpret_cast<unsigned long long*>(&info->out_bytes), (unsigned long long)len);
    if (off + len > out_cap) { *overflow = 1u; return nullptr; }
    return out + off;
}

template <bool SMEM>  // tier 1: the scratch lives in shared memory (known address space: LDS / STS instead of generic accesses)
__global__ void __launch_bounds__(IE_GENERAL_SMALL_THREADS) ie_resolve_general_kernel(const IeTableView* __restrict__ views, uint64_t per_state,
                                                                const uint8_t* __restrict__ tmpl,
                                                                const uint64_t* __restrict__ offs, uint8_t* __restrict__ out,
                                                                uint64_t out_cap, uint64_t* __restrict__ out_offs,
                                                                uint32_t* __restrict__ out_lens, int32_t* __restrict__ status_out,
                                                                uint32_t* __restrict__ aux_out, IeWorkspace ws, ie_batch_info* info,
                                                                uint32_t max_expansions, uint32_t tcap, uint32_t kcap, uint64_t out_bias,
                                                                const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count,
                                                                uint32_t* __restrict__ retry_list, uint32_t* __restrict__ retry_count,
                                                                uint32_t smem_stride, uint32_t fcap) {
    // Two tiers share this kernel: many workers with a small scratch each take the punted templates first
    // (retry_list != nullptr: a template that outgrows the small scratch is queued there, nothing is written for
    // it), then a few workers with the full-size scratch take the queue.
    // ONE TEMPLATE PER WARP, the machine worked by lane 0: it is a byte-serial automaton whose control flow differs
    // per template, so 32 templates on the lanes of one warp run one after the other anyway (measured: 2.5 ms per
    // template that way); one lane per warp lets the SM interleave dozens of independent automata instead.  The other
    // lanes join for the bulk passes around the machine (brace count, result copies).
    // Launched with programmatic stream serialisation (IE_PDL): the launch itself overlaps the tail of the kernel before it
    // in the stream; nothing that kernel wrote (the work lists and their counts) is read before this wait returns.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t worker = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // the full-size tier has ws.general_workers scratch areas, which need not fill the last block
    const uint32_t grid_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_workers = SMEM ? grid_warps : min(grid_warps, ws.general_workers);
    if (worker >= n_workers) return;
    const uint32_t count = *list_count;
    if (worker == 0 && lane == 0 && count && retry_list) atomicAdd(reinterpret_cast<unsigned long long*>(&info->n_general), (unsigned long long)count);
    // The machine re-reads what it just wrote, byte by byte: in global memory every such read is an L2 round trip
    // (stores do not allocate in L1).  Tier 1 therefore keeps its small scratch in SHARED memory (smem_stride != 0,
    // skewed by 4 bytes per thread against bank conflicts); tier 2 uses the big global scratch.
    extern __shared__ __align__(16) uint8_t gen_smem[];
    // per worker: [text tcap][key kcap][frame stack fcap * 16]  (tier 1: text and key in shared memory, MAXF local frames)
    uint8_t* T = SMEM ? gen_smem + (size_t)(threadIdx.x >> 5) * smem_stride
                      : ws.scratch + (size_t)worker * ((size_t)tcap + kcap + (size_t)fcap * sizeof(Frame));
    uint8_t* kscr = T + tcap;
    Frame local_frames[SMEM ? MAXF : 1];
    Frame* const frames = SMEM ? local_frames : reinterpret_cast<Frame*>(kscr + kcap);
    const uint32_t max_frames = SMEM ? MAXF : fcap;

    for (uint32_t q = worker; q < count; q += n_workers) {
        const uint32_t r = list[q];                         // result index = state * per_state + template
        const uint32_t i = (uint32_t)(r % per_state);
        const IeTableView tv = views[r / per_state];
        const uint8_t* t = tmpl + offs[i];
        const uint32_t n = (uint32_t)(offs[i + 1] - offs[i]);

        // peel simple layers (interp.rs:45-52, recursion on the inner key)
        uint32_t lo = 0, hi = n, m = 0;
        if (lane == 0) while (is_simple_range(t, lo, hi)) { ++lo; --hi; ++m; }
        lo = __shfl_sync(FULL, lo, 0); hi = __shfl_sync(FULL, hi, 0); m = __shfl_sync(FULL, m, 0);

        uint32_t nf = 0, ttop = tcap, in_open = 0, in_close = 0, t_close = 0, expansions = 0;
        uint32_t status = IE_RES_STRING, aux = 0;
        const uint8_t* payload = nullptr;  // final bytes when not in T
        uint32_t payload_len = 0;
        bool uneven = false, scratch_full = false;
        if (hi > lo) warp_count_braces(t + lo, hi - lo, lane, in_open, in_close);
        enum { DO_SKIP = 0, DO_UNEVEN, DO_TEXT, DO_PAYLOAD };
        uint32_t todo = DO_SKIP;  // what the warp writes once lane 0 is through
        if (lane == 0) {
        if (hi > lo) frames[nf++] = Frame{t + lo, hi - lo, hi - lo};
        // right-to-left rewriting machine: T holds the (sentinelised) text to the right of the
        // rightmost unresolved '{'; frames hold what is still to its left.
        while (nf > 0 && status == IE_RES_STRING) {
            Frame& f = frames[nf - 1];
            if (f.pos == 0) { --nf; continue; }
            const uint8_t c = f.ptr[f.pos - 1];
            if (c == '{' || c == '}') {
                if (f.pos >= 2 && f.ptr[f.pos - 2] == '\\') {  // escaped: sentinel text
                    if (ttop < 4) { status = IE_RES_LIMIT; scratch_full = true; break; }
                    ttop -= 4;
                    if (c == '{') { T[ttop] = 0x2E; T[ttop + 1] = 0xE3; T[ttop + 2] = 0x80; T[ttop + 3] = 0xA0; }
                    else { T[ttop] = 0xE3; T[ttop + 1] = 0x80; T[ttop + 2] = 0xA0; T[ttop + 3] = 0x2E; }
                    f.pos -= 2;
                    continue;
                }
                if (c == '}') {
                    if (ttop < 1) { status = IE_RES_LIMIT; scratch_full = true; break; }
                    T[--ttop] = '}'; --in_close; ++t_close; --f.pos;
                    continue;
                }
                // rightmost unescaped '{' of the current string: one iteration of interp.rs:54-84
                if (in_open != in_close + t_close) { status = IE_RES_UNEVEN; uneven = true; break; }
                uint32_t idx = ttop;
                while (idx < tcap && T[idx] != '}') ++idx;
                if (idx == tcap) { status = IE_RES_PANIC; break; }
                // The key text is dead once it has been looked up, and restoring the sentinels never lengthens it: a key
                // that does not fit the key buffer is restored in place (full-size tier; the small tier hands it on).
                uint8_t* kbuf = kscr;
                if (unsentinelise<false>(T + ttop, idx - ttop, nullptr) > kcap) {
                    if (retry_list) { status = IE_RES_LIMIT; scratch_full = true; break; }
                    kbuf = T + ttop;
                }
                const uint32_t klen = unsentinelise<true>(T + ttop, idx - ttop, kbuf);
                payload = kbuf; payload_len = klen;
                if (klen == 0) { status = IE_RES_EMPTY_KEY; break; }
                const IeSlot* s = ie_lookup(tv, kbuf, klen);
                if (!s) { status = is_arg_key(kbuf, klen) ? IE_RES_ARG_MISSING : IE_RES_NOT_FOUND; break; }
                if (!tag_splices(IE_SLOT_TAG(s->vl_tf))) { status = IE_RES_UNSUPPORTED; break; }
                payload = nullptr; payload_len = 0;
                ttop = idx + 1; --t_close; --in_open; --f.pos;
                if (++expansions > max_expansions) { status = IE_RES_LIMIT; break; }
                if (f.pos == 0) --nf;
                const uint32_t vlen = IE_SLOT_VLEN(s->vl_tf);
                if (vlen) {
                    if (nf == max_frames) { status = IE_RES_LIMIT; scratch_full = true; break; }  // tier 1: the full-size tier redoes it
                    const uint8_t* v = tv.base + (size_t)s->val_off16 * 16u;
                    frames[nf++] = Frame{v, vlen, vlen};
                    count_braces(v, vlen, in_open, in_close);
                }
                continue;
            }
            if (c == '\\' && f.pos == f.len && ttop < tcap && T[ttop] == '}') {
                // a spliced value ending in '\' escapes the '}' that now follows it (interp.rs:82-83)
                if (ttop < 3) { status = IE_RES_LIMIT; scratch_full = true; break; }
                ttop += 1; ttop -= 4;
                T[ttop] = 0xE3; T[ttop + 1] = 0x80; T[ttop + 2] = 0xA0; T[ttop + 3] = 0x2E;
                --t_close; --f.pos;
                continue;
            }
            if (ttop < 1) { status = IE_RES_LIMIT; scratch_full = true; break; }
            T[--ttop] = c; --f.pos;
        }

        if (scratch_full && retry_list) {  // outgrew the small scratch: the full-size tier redoes it
            retry_list[atomicAdd(retry_count, 1u)] = r;
        } else if (uneven) {
            todo = DO_UNEVEN;
        } else if (status == IE_RES_STRING && m == 0) {
            todo = DO_TEXT;
        } else if (status == IE_RES_STRING) {
            // simple path: the core's string is the first key; each layer looks up the
            // rendering of the previous result, typed and without rescan (interp.rs:47-51)
            uint32_t klen = unsentinelise<false>(T + ttop, tcap - ttop, nullptr);
            const uint8_t* key = kscr;
            if (klen > kcap && retry_list) {
                // the small tier's key buffer is too short: the full-size tier redoes the template
                retry_list[atomicAdd(retry_count, 1u)] = r;
            } else {
                if (klen > kcap) key = T + ttop;  // restored in place, as above
                unsentinelise<true>(T + ttop, tcap - ttop, const_cast<uint8_t*>(key));
                for (uint32_t layer = 0; layer < m; ++layer) {
                    payload = key; payload_len = klen;
                    if (klen == 0) { status = IE_RES_EMPTY_KEY; break; }
                    const IeSlot* s = ie_lookup(tv, key, klen);
                    if (!s) { status = is_arg_key(key, klen) ? IE_RES_ARG_MISSING : IE_RES_NOT_FOUND; break; }
                    key = tv.base + (size_t)s->val_off16 * 16u;
                    klen = IE_SLOT_VLEN(s->vl_tf);
                    payload = key; payload_len = klen;
                    status = IE_RES_TYPED | (IE_SLOT_TAG(s->vl_tf) << 8);
                    aux = s->entry;
                }
                todo = DO_PAYLOAD;
            }
        } else {
            if (status == IE_RES_PANIC || status == IE_RES_LIMIT) payload_len = 0;  // error with key payload (or none)
            todo = DO_PAYLOAD;
        }
        }  // lane 0
        __syncwarp();  // T and the key buffer were written by lane 0
        todo = __shfl_sync(FULL, todo, 0);
        if (todo == DO_SKIP) continue;
        ttop = __shfl_sync(FULL, ttop, 0);
        uint64_t off = 0;
        uint32_t olen = 0;
        uint8_t* w = nullptr;
        if (todo == DO_UNEVEN) {
            // payload = the current string: unread frame prefixes (sentinelised) + T   (interp.rs:57-61)
            nf = __shfl_sync(FULL, nf, 0);
            for (uint32_t k = 0; k < nf; ++k)
                olen += warp_sentinelise(warp_bcast_ptr(frames[k].ptr), __shfl_sync(FULL, frames[k].pos, 0), nullptr, lane);
            olen += tcap - ttop;
            if (lane == 0) w = reserve_out(out, out_cap, info, ws.overflow, olen, off);
            w = warp_bcast_ptr(w);
            if (w) {
                for (uint32_t k = 0; k < nf; ++k)
                    w += warp_sentinelise(warp_bcast_ptr(frames[k].ptr), __shfl_sync(FULL, frames[k].pos, 0), w, lane);
                warp_copy(w, T + ttop, tcap - ttop, lane);
            }
        } else if (todo == DO_TEXT) {
            // restoring the sentinels only touches text that holds their lead byte; without it the result is T as it is
            if (!warp_has_byte(T + ttop, tcap - ttop, 0xE3, lane)) {
                olen = tcap - ttop;
                if (lane == 0) w = reserve_out(out, out_cap, info, ws.overflow, olen, off);
                w = warp_bcast_ptr(w);
                if (w) warp_copy(w, T + ttop, olen, lane);
            } else if (lane == 0) {
                olen = unsentinelise<false>(T + ttop, tcap - ttop, nullptr);
                w = reserve_out(out, out_cap, info, ws.overflow, olen, off);
                if (w) unsentinelise<true>(T + ttop, tcap - ttop, w);
            }
        } else {
            payload = warp_bcast_ptr(payload);
            olen = __shfl_sync(FULL, payload_len, 0);
            if (olen) {
                if (lane == 0) w = reserve_out(out, out_cap, info, ws.overflow, olen, off);
                w = warp_bcast_ptr(w);
                if (w) warp_copy(w, payload, olen, lane);
            }
        }
        __syncwarp();  // the next template's machine overwrites T
        if (lane != 0) continue;
        out_offs[r] = off + out_bias;
        out_lens[r] = olen;
        status_out[r] = (int32_t)status;
        aux_out[r] = aux;
        if ((status & 0xFF) == IE_RES_LIMIT) atomicAdd(reinterpret_cast<unsigned long long*>(&info->n_limit), 1ull);
    }
}

// get_interpdata (interp.rs:91-137) for a batch of literal keys: one thread per key.
__global__ void __launch_bounds__(256) ie_lookup_kernel(IeTableView tv, const uint8_t* __restrict__ keys, const uint64_t* __restrict__ offs,
                                                        uint64_t n, int32_t* __restrict__ tag_out, uint32_t* __restrict__ entry_out) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t a = offs[k];
    const uint32_t len = (uint32_t)(offs[k + 1] - a);
    const IeSlot* s = len ? ie_lookup(tv, keys + a, len) : nullptr;
    tag_out[k] = s ? (int32_t)IE_SLOT_TAG(s->vl_tf) : -1;
    entry_out[k] = s ? s->entry : IE_AUX_NONE;
}

// ---- rescan rounds: plumbing between two launches of the tile kernel --------------------------------------------
// Claims room for the next round's template arena behind the results (same bump allocator as the tiles) and resets the
// counters the round will write.  `in` = which again list the round reads.
__global__ void ie_round_reserve_kernel(IeRoundCtl* ctl, uint32_t in, ie_batch_info* info, uint64_t out_cap, uint32_t* overflow) {
    const uint64_t bytes = ctl->bytes[in];
    const uint64_t base = atomicAdd(reinterpret_cast<unsigned long long*>(&info->out_bytes), (unsigned long long)((bytes + 15) & ~15ull));
    if (base + bytes > out_cap) { *overflow = 1u; ctl->count[in] = 0; }  // the caller regrows the arena and reruns the batch
    ctl->base = base;
    ctl->packed = 0;
    ctl->count[in ^ 1] = 0;
    ctl->bytes[in ^ 1] = 0;
}
// One warp per unfinished template: its text so far (a result of the previous round) is copied to the round's arena.
// One packed atomic hands out (index, byte offset) together, so the texts are contiguous in index order - the layout
// the tile kernel streams.
__global__ void __launch_bounds__(256) ie_round_gather_kernel(IeRoundCtl* ctl, uint32_t in, const uint32_t* __restrict__ list,
                                                              uint8_t* __restrict__ out, const uint64_t* __restrict__ out_offs,
                                                              const uint32_t* __restrict__ out_lens, uint64_t out_bias,
                                                              uint64_t* __restrict__ offs2, uint32_t* __restrict__ map2) {
    const uint32_t count = ctl->count[in];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < count; q += warps) {
        const uint32_t entry = list[q];
        const uint32_t r = entry & IE_AGAIN_INDEX_MASK;
        const uint32_t len = out_lens[r];
        const uint8_t* src = out + (out_offs[r] - out_bias);
        unsigned long long old = 0;
        if (lane == 0) old = atomicAdd(reinterpret_cast<unsigned long long*>(&ctl->packed), (1ull << 40) | (unsigned long long)len);
        old = __shfl_sync(0xFFFFFFFFu, old, 0);
        const uint32_t idx = (uint32_t)(old >> 40);
        const uint64_t at = ctl->base + (old & ((1ull << 40) - 1));
        if (lane == 0) {
            offs2[idx] = at;
            map2[idx] = entry;
            if (idx + 1 == count) offs2[count] = at + len;
        }
        uint8_t* dst = out + at;
        for (uint32_t k = lane; k < len; k += 32) dst[k] = src[k];
    }
}

}  // namespace

cudaError_t ie_launch_lookup(const IeTableView& tv, const uint8_t* d_keys, const uint64_t* d_offs, uint64_t n, int32_t* d_tag,
                             uint32_t* d_entry, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    ie_lookup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tv, d_keys, d_offs, n, d_tag, d_entry);
    return cudaGetLastError();
}

cudaError_t ie_launch_resolve(const IeTableView* d_views, uint32_t n_states, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                              uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                              const IeWorkspace& ws, ie_batch_info* d_info, uint32_t max_expansions, uint32_t tcap,
                              uint64_t out_bias, uint32_t tt, uint32_t rescan_rounds, uint64_t table_bytes, cudaStream_t stream) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(ws.zero_base, 0, ws.zero_bytes, stream)) != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(d_info, 0, sizeof(ie_batch_info), stream)) != cudaSuccess) return err;
    if (n == 0) return cudaSuccess;
    // rounds need one snapshot (the table must stay tile-uniform) and result indices that fit the round map
    // (the gather hands out index and byte offset in one 64-bit atomic, 24 bits of it for the index)
    if (!ws.round_ctl || n * n_states >= (1u << 24)) rescan_rounds = 0;
    IeRound rd{};
    rd.allow_splice = rescan_rounds ? 1u : 0u;
    if (rescan_rounds) { rd.again_list = ws.round_list[0]; rd.again_count = &ws.round_ctl->count[0]; rd.again_bytes = &ws.round_ctl->bytes[0]; }
    // many snapshots with a handful of templates each: 32-template tiles on 64-thread CTAs (a tile never mixes snapshots)
    const bool small_tiles = n_states > 1 && n <= IE_SMALL_TILE;
    if (small_tiles) err = ie_launch_resolve_tiles_small(d_views, n_states, d_tmpl, d_offs, n, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, ws,
                                                         d_info, out_bias, tt < IE_SMALL_TILE ? tt : IE_SMALL_TILE, rd, stream);
#ifndef IE_NO_FUSED
    // The fused single-pass kernel takes every launch without rescan rounds (the rounds' splice bookkeeping lives in the
    // phase-wise kernel) on a table whose value references fit 31 bits of 16-byte units.
    else if (!rescan_rounds && table_bytes < (1ull << 35))
        err = ie_launch_resolve_fused(d_views, n_states, d_tmpl, d_offs, n, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, ws,
                                                           d_info, out_bias, tt, stream);
#endif
    else err = ie_launch_resolve_tiles(d_views, n_states, d_tmpl, d_offs, n, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, ws, d_info,
                                       out_bias, tt, rd, stream);
    if (err != cudaSuccess) return err;
    for (uint32_t k = 0; k < rescan_rounds; ++k) {
        // round k + 2 reads again list (k & 1), maps results through list 2 and fills again list ((k + 1) & 1)
        const uint32_t in = k & 1;
        ie_round_reserve_kernel<<<1, 1, 0, stream>>>(ws.round_ctl, in, d_info, out_cap, ws.overflow);
        ie_round_gather_kernel<<<296, 256, 0, stream>>>(ws.round_ctl, in, ws.round_list[in], d_out, d_out_offs, d_out_lens, out_bias, ws.round_offs,
                                                        ws.round_list[2]);
        if ((err = cudaGetLastError()) != cudaSuccess) return err;
        IeRound r2{};
        r2.n_dev = &ws.round_ctl->count[in];
        r2.bytes_dev = &ws.round_ctl->bytes[in];
        r2.result_map = ws.round_list[2];
        r2.again_list = ws.round_list[in ^ 1];
        r2.again_count = &ws.round_ctl->count[in ^ 1];
        r2.again_bytes = &ws.round_ctl->bytes[in ^ 1];
        r2.allow_splice = 1;
        r2.last_round = k + 1 == rescan_rounds;
        r2.per_state = n_states > 1 ? (uint32_t)n : 0u;  // several snapshots: every template of the round finds its own table
        r2.views_all = d_views;
        // templates = the gathered texts inside the out arena (absolute offsets in round_offs); results still go to
        // the caller's arrays at the original result indices
        if ((err = ie_launch_resolve_tiles(d_views, 1, d_out, ws.round_offs, n * n_states, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, ws,
                                           d_info, out_bias, tt, r2, stream)) != cudaSuccess)
            return err;
    }
    // tier 1: three blocks of 16 warps per SM, one template per warp, IE_GENERAL_SMALL_TEXT + IE_GENERAL_SMALL_KEY bytes
    // of shared memory per warp; tier 2: ws.general_workers warps with the caller's full limits in global memory
    const uint32_t small_t = tcap < IE_GENERAL_SMALL_TEXT ? tcap : IE_GENERAL_SMALL_TEXT;
    const uint32_t stride = IE_GENERAL_SMALL_TEXT + IE_GENERAL_SMALL_KEY + 4;
    const size_t smem = (size_t)(IE_GENERAL_SMALL_THREADS / 32) * stride;
    if ((err = cudaFuncSetAttribute(ie_resolve_general_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // Both tiers are usually empty (C4: nothing is punted) and then cost only their launch; with programmatic stream
    // serialisation that launch overlaps the tail of the kernel in front (the kernels start with griddepcontrol.wait).
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.stream = stream;
    cfg.attrs = pdl;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(sms * 3));
    cfg.blockDim = dim3(IE_GENERAL_SMALL_THREADS);
    cfg.dynamicSmemBytes = smem;
    const uint64_t per_state = n;
    const uint32_t key_small = IE_GENERAL_SMALL_KEY, key_full = IE_KEY_SCRATCH, zero = 0u, fcap = ie_general_fcap(tcap);
    uint32_t* const no_list = nullptr;
    if ((err = cudaLaunchKernelEx(&cfg, ie_resolve_general_kernel<true>, d_views, per_state, d_tmpl, d_offs, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, ws,
                                  d_info, max_expansions, small_t, key_small, out_bias, (const uint32_t*)ws.general_list, (const uint32_t*)ws.general_count,
                                  ws.retry_list, ws.retry_count, stride, zero)) != cudaSuccess)
        return err;
    cfg.gridDim = dim3(ws.general_workers / (IE_GENERAL_SMALL_THREADS / 32));
    cfg.dynamicSmemBytes = 0;
    if ((err = cudaLaunchKernelEx(&cfg, ie_resolve_general_kernel<false>, d_views, per_state, d_tmpl, d_offs, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, ws,
                                  d_info, max_expansions, tcap, key_full, out_bias, (const uint32_t*)ws.retry_list, (const uint32_t*)ws.retry_count, no_list,
                                  no_list, zero, fcap)) != cudaSuccess)
        return err;
    return cudaGetLastError();
}

// Re-runs the templates on `d_list` (result indices, *d_count of them) on the full-size tier with larger bounds: the
// host-buffer calls escalate templates that hit a DEFAULT bound (ie_capi.cu: escalate_limits).  ws.scratch must hold
// ws.general_workers * ie_general_worker_bytes(tcap); results are appended to d_out through d_info->out_bytes.
cudaError_t ie_launch_general_escalate(const IeTableView* d_views, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                                       uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                                       const IeWorkspace& ws, ie_batch_info* d_info, uint32_t max_expansions, uint32_t tcap, uint64_t out_bias,
                                       const uint32_t* d_list, const uint32_t* d_count, cudaStream_t stream) {
    const uint32_t wpb = IE_GENERAL_SMALL_THREADS / 32;
    ie_resolve_general_kernel<false><<<(ws.general_workers + wpb - 1) / wpb, IE_GENERAL_SMALL_THREADS, 0, stream>>>(
        d_views, n, d_tmpl, d_offs, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, ws, d_info, max_expansions, tcap, IE_KEY_SCRATCH, out_bias,
        d_list, d_count, nullptr, nullptr, 0u, ie_general_fcap(tcap));
    return cudaGetLastError();
}
