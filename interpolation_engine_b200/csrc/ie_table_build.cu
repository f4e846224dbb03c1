// Device-side construction and in-place mutation of the packed inserts table (sm_100a).
//
// Replaces, for the snapshot argument of every interp.rs function (`&Map<String, Value>`, cloned per task at
// rust-project/src/runtime.rs:700) and for set_interpdata / delete_interpdata (interp.rs:139-145):
//
//   ie_table_claim_kernel / ie_table_fill_kernel   build the open-addressing tables of MANY snapshots in one pair of
//       launches from the raw packed arrays (key arena, value arena, tags, snapshot offsets): one thread per insert
//       hashes its key (murmur3-32), claims a slot with atomicCAS, classifies its value (ie_classify_value) and copies key
//       and value into the slot's inline areas or into the arena behind the slots (one atomic bump per long item).
//       The host does O(snapshots) work: capacities and the layout; nothing per insert.
//   ie_table_mutate_kernel   applies an ordered list of set / delete operations to one snapshot (or to every snapshot
//       of the table) in place: overwrite, insert into the first tombstone / empty slot of the probe chain, tombstone.
//
// Layout of a device-built table (one allocation; every reference inside snapshot s is in 16-byte units from ITS slot
// array, so the resolve kernels see nothing new):
//
//   [slots of snapshot 0][slots of snapshot 1]...[arena: long keys and values, bump-allocated][slack][IeTableView[S]][IeTableHeader]
#include <cuda_runtime.h>

#include "ie_common.cuh"
#include "ie_kernels.h"

namespace {

constexpr uint32_t OWNER_NONE = 0xFFFFFFFFu;  // what the 0xFF memset leaves in a slot's pad[0]

__device__ __forceinline__ uint64_t pad16(uint64_t x) { return (x + 15) & ~uint64_t(15); }

// The packed arrays of a build as every insert sees them: item g in [0, n) is a caller insert, the two items after the
// inserts of each snapshot are its clock keys "HH:MM" / "HH:MM:SS" (interp.rs:96-104: answered before the map, so they
// are inserted last and win over inserts of the same name).
struct Item {
    const uint8_t* key; uint32_t key_len;
    const uint8_t* val; uint32_t val_len;
    uint32_t tag, entry, state;
    bool valid;
};

__device__ __forceinline__ Item load_item(const IeBuildArgs& a, uint64_t g) {
    Item it{};
    const uint64_t n_total = a.n + (a.with_clock ? 2ull * a.n_states : 0ull);
    it.valid = g < n_total;
    if (!it.valid) return it;
    if (g < a.n) {
        // snapshot of insert g: the last s with state_offs[s] <= g
        uint32_t lo = 0, hi = a.n_states;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.state_offs[mid] <= g) lo = mid; else hi = mid; }
        it.state = lo;
        const uint64_t k0 = a.key_offs[g], k1 = a.key_offs[g + 1], v0 = a.val_offs[g], v1 = a.val_offs[g + 1];
        it.key = a.keys + k0; it.key_len = (uint32_t)(k1 - k0);
        it.val = a.vals + v0; it.val_len = (uint32_t)(v1 - v0);
        it.tag = a.tags[g];
        it.entry = (uint32_t)(g - a.state_offs[lo]);
        if (k1 < k0 || v1 < v0 || k1 - k0 >= IE_SLOT_TOMB || v1 - v0 > IE_VLEN_MAX || it.tag > IE_TAG_OBJECT) { atomicOr(&a.hdr->error, 1u); it.valid = false; }
    } else {
        const uint64_t c = g - a.n;
        it.state = (uint32_t)(c >> 1);
        const bool sec = c & 1;
        if (!((a.with_clock >> (sec ? 1 : 0)) & 1u)) { it.valid = false; return it; }
        it.key = a.clock + (sec ? 8 : 0);
        it.key_len = sec ? 8 : 5;
        it.val = a.clock + (sec ? 80 : 16);
        it.val_len = sec ? a.hhmmss_len : a.hhmm_len;
        it.tag = IE_TAG_STRING;
        it.entry = (uint32_t)(a.state_offs[it.state + 1] - a.state_offs[it.state]) + (sec ? 1u : 0u);
    }
    return it;
}

__device__ __forceinline__ bool same_key(const Item& a, const Item& b) {
    if (a.key_len != b.key_len) return false;
    for (uint32_t i = 0; i < a.key_len; ++i) if (a.key[i] != b.key[i]) return false;
    return true;
}

// Pass 1: every item claims the slot of its key.  The claim word (pad[0] of the slot) holds the global index of the
// item that owns the slot; it only ever changes to a LATER item with the same key (Map::insert: later duplicates win),
// so a prober can always compare against the owner's key in the input arrays - no locks, no spinning.
__global__ void __launch_bounds__(256) ie_table_claim_kernel(IeBuildArgs a) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const Item it = load_item(a, g);
    if (!it.valid) return;
    const uint32_t h = ie_hash_bytes(it.key, it.key_len);
    IeSlot* slots = reinterpret_cast<IeSlot*>(a.base + a.slot_base[it.state] * sizeof(IeSlot));
    const uint32_t mask = a.slot_cap[it.state] - 1;
    uint32_t idx = h & mask;
    for (uint32_t probes = 0;; ++probes) {
        uint32_t* owner = &slots[idx].pad[0];
        uint32_t cur = atomicCAS(owner, OWNER_NONE, (uint32_t)g);
        if (cur == OWNER_NONE) break;  // fresh slot: mine
        if (same_key(it, load_item(a, cur))) { atomicMax(owner, (uint32_t)g); break; }
        idx = (idx + 1) & mask;
        if (probes > mask) { atomicOr(&a.hdr->error, 2u); return; }  // cannot happen at load factor <= 0.5
    }
    a.slot_of[g] = idx;
}

// Pass 2: the owner of each slot writes it.
__global__ void __launch_bounds__(256) ie_table_fill_kernel(IeBuildArgs a) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const Item it = load_item(a, g);
    if (!it.valid) return;
    uint8_t* sbase = a.base + a.slot_base[it.state] * sizeof(IeSlot);
    IeSlot* s = reinterpret_cast<IeSlot*>(sbase) + a.slot_of[g];
    if (s->pad[0] != (uint32_t)g) return;  // a later insert of the same key won
    uint8_t* arena = a.base + a.arena_off;
    s->hash = ie_hash_bytes(it.key, it.key_len);
    s->entry = it.entry;
    uint8_t* kdst = s->key_inline;
    if (it.key_len > IE_INLINE_BYTES) {
        const uint64_t at = atomicAdd(reinterpret_cast<unsigned long long*>(&a.hdr->arena_used), (unsigned long long)pad16(it.key_len));
        if (at + pad16(it.key_len) > a.arena_bytes) { atomicOr(&a.hdr->error, 4u); return; }
        kdst = arena + at;
    }
    for (uint32_t i = 0; i < IE_INLINE_BYTES; ++i) s->key_inline[i] = 0;
    for (uint32_t i = 0; i < it.key_len; ++i) kdst[i] = it.key[i];
    s->key_off16 = (uint32_t)((kdst - sbase) >> 4);
    uint8_t* vdst = s->val_inline;
    if (it.val_len > IE_INLINE_BYTES) {
        const uint64_t at = atomicAdd(reinterpret_cast<unsigned long long*>(&a.hdr->arena_used), (unsigned long long)pad16(it.val_len));
        if (at + pad16(it.val_len) > a.arena_bytes) { atomicOr(&a.hdr->error, 4u); return; }
        vdst = arena + at;
    }
    for (uint32_t i = 0; i < IE_INLINE_BYTES; ++i) s->val_inline[i] = 0;
    for (uint32_t i = 0; i < it.val_len; ++i) vdst[i] = it.val[i];
    s->val_off16 = (uint32_t)((vdst - sbase) >> 4);
    const uint32_t vflags = ie_classify_value(it.val, it.val_len);
    if (vflags & IE_VF_BALANCED) atomicOr(&a.hdr->flags, 1u);
    s->vl_tf = it.val_len | (it.tag << 25) | (vflags << 28);
    atomicAdd(&a.used[it.state], 1u);
    __threadfence();
    s->key_len = it.key_len;  // last: the slot is live
}

__global__ void ie_table_views_kernel(IeBuildArgs a, IeTableView* views) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_states) return;
    views[s].base = a.base + a.slot_base[s] * sizeof(IeSlot);
    views[s].mask = a.slot_cap[s] - 1;
    views[s].n_entries = (uint32_t)(a.state_offs[s + 1] - a.state_offs[s]);
}

// ---- set_interpdata / delete_interpdata in place (interp.rs:139-145) --------------------------------------------------
// One warp per target snapshot; lane 0 walks the operations in order (a later operation on the same key sees the
// earlier one), the warp copies the bytes.  A value that no longer fits where the old one lives takes fresh arena space
// (the old space is reclaimed when the table is packed again).
__global__ void __launch_bounds__(32) ie_table_mutate_kernel(IeMutateArgs m) {
    const uint32_t lane = threadIdx.x;
    const uint32_t state = m.all_states ? blockIdx.x : m.state;
    const IeTableView tv = m.views[state];
    uint8_t* sbase = const_cast<uint8_t*>(tv.base);
    IeSlot* slots = reinterpret_cast<IeSlot*>(sbase);
    uint8_t* arena = m.base + m.arena_off;
    for (uint32_t op = 0; op < m.n_ops; ++op) {
        const uint64_t k0 = m.key_offs[op], k1 = m.key_offs[op + 1];
        const uint8_t* key = m.keys + k0;
        const uint32_t klen = (uint32_t)(k1 - k0);
        const bool is_set = m.vals != nullptr;
        // lane 0: find the key's slot, or where it would go
        uint32_t found = 0xFFFFFFFFu, place = 0xFFFFFFFFu;
        if (lane == 0) {
            const uint32_t h = ie_hash_bytes(key, klen);
            uint32_t idx = h & tv.mask;
            for (uint32_t probes = 0; probes <= tv.mask; ++probes) {
                const IeSlot& s = slots[idx];
                if (s.key_len == IE_SLOT_EMPTY) { if (place == 0xFFFFFFFFu) place = idx; break; }
                if (s.key_len == IE_SLOT_TOMB) { if (place == 0xFFFFFFFFu) place = idx; }
                else if (s.hash == h && s.key_len == klen) {
                    const uint8_t* stored = sbase + (size_t)s.key_off16 * 16u;
                    uint32_t i = 0;
                    for (; i < klen; ++i) if (stored[i] != key[i]) break;
                    if (i == klen) { found = idx; break; }
                }
                idx = (idx + 1) & tv.mask;
            }
        }
        found = __shfl_sync(0xFFFFFFFFu, found, 0);
        place = __shfl_sync(0xFFFFFFFFu, place, 0);
        if (!is_set) {  // delete_interpdata: a tombstone keeps the probe chains through this slot intact
            if (found != 0xFFFFFFFFu && lane == 0) { slots[found].hash = 0; slots[found].key_len = IE_SLOT_TOMB; }
            __syncwarp();
            continue;
        }
        const uint64_t v0 = m.val_offs[op], v1 = m.val_offs[op + 1];
        const uint8_t* val = m.vals + v0;
        const uint32_t vlen = (uint32_t)(v1 - v0);
        uint32_t target = found;
        if (found == 0xFFFFFFFFu) {  // new key
            // keep one slot in four free: an insert into a table that is fuller than that is refused (the caller re-packs)
            uint32_t ok = 1;
            if (lane == 0) {
                if (place == 0xFFFFFFFFu) ok = 0;
                else if (slots[place].key_len == IE_SLOT_EMPTY && m.used_slots[state] + 1 > (tv.mask + 1) - (tv.mask + 1) / 4) ok = 0;
            }
            ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
            if (!ok) { if (lane == 0) atomicOr(&m.hdr->error, 8u); continue; }
            target = place;
        }
        IeSlot* s = slots + target;
        // where the value goes: inline, the old arena place when it is large enough, or fresh arena space
        uint8_t* vdst = s->val_inline;
        uint32_t fail = 0;
        if (vlen > IE_INLINE_BYTES) {
            unsigned long long at = 0;
            bool reuse = false;
            if (lane == 0) {
                if (found != 0xFFFFFFFFu) {
                    const uint32_t old_len = IE_SLOT_VLEN(s->vl_tf);
                    reuse = old_len > IE_INLINE_BYTES && pad16(old_len) >= pad16(vlen);
                    if (reuse) at = (unsigned long long)(sbase + (size_t)s->val_off16 * 16u - arena);
                }
                if (!reuse) {
                    at = atomicAdd(reinterpret_cast<unsigned long long*>(&m.hdr->arena_used), (unsigned long long)pad16(vlen));
                    if (at + pad16(vlen) > m.arena_bytes) fail = 1;
                }
            }
            at = __shfl_sync(0xFFFFFFFFu, at, 0);
            fail = __shfl_sync(0xFFFFFFFFu, fail, 0);
            if (fail) { if (lane == 0) atomicOr(&m.hdr->error, 4u); continue; }
            vdst = arena + at;
        }
        uint8_t* kdst = s->key_inline;
        if (found == 0xFFFFFFFFu && klen > IE_INLINE_BYTES) {
            unsigned long long at = 0;
            if (lane == 0) {
                at = atomicAdd(reinterpret_cast<unsigned long long*>(&m.hdr->arena_used), (unsigned long long)pad16(klen));
                if (at + pad16(klen) > m.arena_bytes) fail = 1;
            }
            at = __shfl_sync(0xFFFFFFFFu, at, 0);
            fail = __shfl_sync(0xFFFFFFFFu, fail, 0);
            if (fail) { if (lane == 0) atomicOr(&m.hdr->error, 4u); continue; }
            kdst = arena + at;
        }
        // The slot is taken out of service while its bytes change (no resolve runs concurrently on this engine's
        // stream, but a half-written slot must never look live).
        const uint32_t was_empty = (found == 0xFFFFFFFFu && s->key_len == IE_SLOT_EMPTY) ? 1u : 0u;
        __syncwarp();
        if (lane == 0) s->key_len = IE_SLOT_TOMB;
        __syncwarp();
        if (found == 0xFFFFFFFFu) {
            if (lane < IE_INLINE_BYTES) s->key_inline[lane] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < klen; i += 32) kdst[i] = key[i];
        }
        if (lane < IE_INLINE_BYTES) s->val_inline[lane] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < vlen; i += 32) vdst[i] = val[i];
        __syncwarp();
        if (lane == 0) {
            if (found == 0xFFFFFFFFu) {
                s->hash = ie_hash_bytes(key, klen);
                s->key_off16 = (uint32_t)((kdst - sbase) >> 4);
                if (was_empty) atomicAdd(&m.used_slots[state], 1u);
            }
            s->val_off16 = (uint32_t)((vdst - sbase) >> 4);
            s->entry = m.entries ? m.entries[op] : IE_AUX_NONE;
            const uint32_t vflags = m.flags[op];
            if (vflags & IE_VF_BALANCED) atomicOr(&m.hdr->flags, 1u);
            s->vl_tf = vlen | ((uint32_t)m.tags[op] << 25) | (vflags << 28);
            __threadfence();
            s->key_len = klen;
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t ie_launch_table_build(const IeBuildArgs& a, IeTableView* d_views, cudaStream_t stream) {
    const uint64_t items = a.n + (a.with_clock ? 2ull * a.n_states : 0ull);
    ie_table_views_kernel<<<(a.n_states + 255) / 256, 256, 0, stream>>>(a, d_views);
    if (items) {
        const unsigned blocks = (unsigned)((items + 255) / 256);
        ie_table_claim_kernel<<<blocks, 256, 0, stream>>>(a);
        ie_table_fill_kernel<<<blocks, 256, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t ie_launch_table_mutate(const IeMutateArgs& m, uint32_t n_states, cudaStream_t stream) {
    if (!m.n_ops) return cudaSuccess;
    ie_table_mutate_kernel<<<m.all_states ? n_states : 1u, 32, 0, stream>>>(m);
    return cudaGetLastError();
}
