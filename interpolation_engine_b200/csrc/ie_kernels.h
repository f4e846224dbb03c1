// Internal launch interface between the C ABI (ie_capi.cpp) and the CUDA translation units.
#pragma once
#include <cuda_runtime.h>

#include "ie_common.cuh"

#ifndef IE_RESOLVE_TILE
#define IE_RESOLVE_TILE 128   // templates per CTA tile of the resolve kernel at most
#endif
#ifndef IE_M_PER
#define IE_M_PER 17              // 16-byte chunks of template text per template a tile can hold (chunk-mask table)
#endif
#define IE_TILE_TEXT_BYTES (IE_M_PER * IE_RESOLVE_TILE * 16u - 864u)  // longer tiles take the per-thread path
#define IE_KEY_SCRATCH 4096u  // key buffer of the general path's full-size tier (longer keys are restored in place)
#define IE_GENERAL_WORKERS 2048u       // tier 2 of the general path: full-size scratch (tcap + IE_KEY_SCRATCH each)
#define IE_GENERAL_SMALL_THREADS 512     // general path: 16 warps per block, one template per warp (tier 1: scratch in shared memory)
#define IE_GENERAL_SMALL_TEXT 1792u
#define IE_GENERAL_SMALL_KEY 256u

// Frame-stack entries (16 bytes each) of a full-size general-path worker with a text scratch of tcap bytes, and the
// worker's whole scratch: [text tcap][key IE_KEY_SCRATCH][frames].
inline uint32_t ie_general_fcap(uint32_t tcap) { return tcap / 64 + 64; }
inline size_t ie_general_worker_bytes(uint32_t tcap) { return (size_t)tcap + IE_KEY_SCRATCH + (size_t)ie_general_fcap(tcap) * 16; }

// Per-engine device workspace.  [zero_base, zero_base + zero_bytes) is cleared before a batch.
struct IeWorkspace {
    uint8_t* zero_base;
    size_t zero_bytes;
    uint32_t* tile_counter;   // dynamic tile ids (look-back forward progress)
    uint32_t* general_count;  // templates handed to the general kernel
    uint32_t* overflow;       // set when the out arena was too small
    uint32_t* fix_count;      // escape kernel: entries of fix_list (may exceed IE_ESCAPE_FIX_CAP)
    uint64_t* fix_list;       // escape kernel: [IE_ESCAPE_FIX_CAP] positions of string-final backslashes
    uint64_t* tile_first;     // escape kernel: [tiles + 1] first string starting at or after each tile
    uint64_t* tile_state;     // [tiles] flag << 62 | bytes
    uint32_t* general_list;   // [n]
    uint32_t* retry_list;     // [n] templates that outgrew the small scratch of the general path's first tier
    uint32_t* retry_count;
    // rescan rounds (nullptr when rounds are off): control block in the zeroed region, three index lists of n
    // entries (two alternate as again lists, one is the current round's result map), offsets of the round's templates
    struct IeRoundCtl* round_ctl;
    uint32_t* round_list[3];
    uint64_t* round_offs;     // [n + 1]
    uint8_t* scratch;         // general_workers * ie_general_worker_bytes(tcap)
    uint32_t general_workers;
};

// Resolves the n templates against each of the n_states packed snapshots (d_views[s], device array): result
// index = s * n + template; every output array holds n_states * n entries.  table_bytes = size of the table allocation.
cudaError_t ie_launch_resolve(const IeTableView* d_views, uint32_t n_states, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                              uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                              const IeWorkspace& ws, ie_batch_info* d_info, uint32_t max_expansions, uint32_t tcap,
                              uint64_t out_bias, uint32_t tt, uint32_t rescan_rounds, uint64_t table_bytes, cudaStream_t stream);

cudaError_t ie_launch_general_escalate(const IeTableView* d_views, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                                       uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                                       const IeWorkspace& ws, ie_batch_info* d_info, uint32_t max_expansions, uint32_t tcap, uint64_t out_bias,
                                       const uint32_t* d_list, const uint32_t* d_count, cudaStream_t stream);

// Rescan rounds (interp.rs:81-83 rescans every spliced value): a template whose lookups returned values with properly
// nested groups of their own is written out with those values in place ("spliced") and resolved again as a template
// of the next round, on the same tile kernel.  A round's inputs are gathered contiguously behind the results in the
// out arena; again_list / result_map entries = result index | simple-path layers of the caller's text << 28.
#define IE_AGAIN_INDEX_MASK 0x0FFFFFFFu
#define IE_AGAIN_LAYER_SHIFT 28
#define IE_AGAIN_LAYER_MAX 14u
#define IE_RES_AGAIN 0xFE  // internal status while a batch is in flight (like IE_RES_PUNT)
#define IE_ROUND_MIN_TILE 32u  // rounds >= 2 are launched with one CTA per 32 templates; the kernel uses fewer, larger tiles when the texts are short
struct IeRoundCtl {       // lives in the workspace's zeroed region
    uint32_t count[2];    // again-list lengths, alternating per round
    uint32_t pad[2];
    uint64_t bytes[2];    // bytes of the texts of the templates on each list
    uint64_t base;        // where the current round's template arena starts in the out arena
    uint64_t packed;      // gather: templates placed << 40 | bytes placed
};
struct IeRound {
    const uint32_t* n_dev;       // rounds >= 2: number of templates (device counter); nullptr in round 1
    const uint64_t* bytes_dev;   // rounds >= 2: bytes of their texts (the kernel sizes its tiles from the mean length)
    const uint32_t* result_map;  // rounds >= 2: [n] result index | layers << 28
    uint32_t* again_list;        // out: templates that need another round
    uint32_t* again_count;
    uint64_t* again_bytes;       // out: bytes of their texts, each rounded up to 16
    uint32_t allow_splice;       // 0: values with groups of their own are punted to the general path (no rounds)
    uint32_t last_round;         // 1: what would need yet another round goes to the general path instead
    // rounds >= 2 on a table of several snapshots: a round's tile mixes templates of different snapshots, so every
    // template finds its own table through its result index (result_map[i] / per_state); 0 = one snapshot
    uint32_t per_state;
    const IeTableView* views_all;
};

cudaError_t ie_launch_resolve_tiles(const IeTableView* d_views, uint32_t n_states, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                                    uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                                    const IeWorkspace& ws, ie_batch_info* d_info, uint64_t out_bias, uint32_t tt, const IeRound& rd,
                                    cudaStream_t stream);

// The fused single-pass tile kernel (ie_resolve_fused.cu): launches without rescan rounds on full-size tiles.
cudaError_t ie_launch_resolve_fused(const IeTableView* d_views, uint32_t n_states, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n, uint8_t* d_out,
                                    uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                                    const IeWorkspace& ws, ie_batch_info* d_info, uint64_t out_bias, uint32_t tt, cudaStream_t stream);

// The same kernel compiled with 32-template tiles / 64-thread CTAs (ie_resolve_tile.cu with IE_TILE_SMALL): used when
// every snapshot brings at most IE_SMALL_TILE templates, tt <= IE_SMALL_TILE.
#define IE_SMALL_TILE 32u
cudaError_t ie_launch_resolve_tiles_small(const IeTableView* d_views, uint32_t n_states, const uint8_t* d_tmpl, const uint64_t* d_offs, uint64_t n,
                                          uint8_t* d_out, uint64_t out_cap, uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status,
                                          uint32_t* d_aux, const IeWorkspace& ws, ie_batch_info* d_info, uint64_t out_bias, uint32_t tt,
                                          const IeRound& rd, cudaStream_t stream);

// Templates per tile for a batch whose templates average `avg_bytes` bytes and `avg_groups_x16` / 16 `{...}` groups
// (0 = unknown, assume short / few): the largest power of two <= IE_RESOLVE_TILE whose expected text, brace events
// (2 per group) and copy segments (2 per group + 1) fit the tile's tables with 20-25 % headroom.  A tile that outgrows
// its tables still resolves exactly, but on the slow per-thread path: this keeps dense templates off it.
#define IE_TILE_EVENTS (23u * 64u)     // ie_resolve_tile.cu: E_CAP of the 128-template build
#define IE_TILE_SEGMENTS (8u * 128u)   // S_CAP
inline uint32_t ie_pick_tile(uint64_t avg_bytes, uint64_t avg_groups_x16 = 0) {
    uint32_t tt = IE_RESOLVE_TILE;
    while (tt > 4 && ((uint64_t)tt * avg_bytes * 5 / 4 > IE_TILE_TEXT_BYTES ||
                      (uint64_t)tt * avg_groups_x16 * 2 * 5 / 4 > 16ull * IE_TILE_EVENTS ||
                      (uint64_t)tt * (avg_groups_x16 * 2 + 16) * 5 / 4 > 16ull * IE_TILE_SEGMENTS))
        tt >>= 1;
    return tt;
}

// tag_out[i] = value tag or -1 on a miss; entry_out[i] = insert index (>= n_entries: clock key)
cudaError_t ie_launch_lookup(const IeTableView& tv, const uint8_t* d_keys, const uint64_t* d_offs, uint64_t n, int32_t* d_tag,
                             uint32_t* d_entry, cudaStream_t stream);

// mode 0 unescape, 1 escape (interp.rs:147-177).  in_bytes >= d_in_offs[n] - d_in_offs[0]; the workspace holds
// ie_escape_tiles(in_bytes) + 1 zeroed look-back words.
#define IE_ESCAPE_TILE_BYTES 8192u
#define IE_ESCAPE_FIX_CAP 1024u
inline uint64_t ie_escape_tiles(uint64_t in_bytes) { return (in_bytes + 15 + IE_ESCAPE_TILE_BYTES - 1) / IE_ESCAPE_TILE_BYTES + 1; }
cudaError_t ie_launch_escape(int mode, const uint8_t* d_in, const uint64_t* d_in_offs, uint64_t n, uint64_t in_bytes, uint8_t* d_out,
                             uint64_t out_cap, uint64_t* d_out_offs, const IeWorkspace& ws, cudaStream_t stream);

#define IE_GLOB_GENERIC 0
#define IE_GLOB_FAST 1  // prefix * suffix (either may be empty; no star at all: exact)
#define IE_GLOB_MID 2   // prefix * middle * suffix: the middle piece is searched between the two
struct IeGlobFast {  // a pattern with at most two '*' runs and literal pieces of <= 32 bytes, compiled by ie_glob_compile
    uint32_t pre[8], pre_mask[8];  // image / mask of the key's first 32 bytes
    uint32_t suf[8], suf_mask[8];  // image / mask of the key's last 32 bytes (right aligned)
    uint16_t min_len;              // prefix + suffix bytes
    uint8_t kind, exact;           // exact: no star, the key must be exactly min_len long
    uint8_t pre_words, suf_first;  // words [0, pre_words) of pre and [suf_first, 8) of suf carry mask bits
    uint8_t probe, complete;       // probe: the prefix word tested first; complete: the probe already is the whole test
    uint32_t mid, mid_mask;        // IE_GLOB_MID: image / mask of the first (up to 4) bytes of the middle piece
    uint16_t mid_len;
    uint8_t mid_lo, mid_hi;        // allowed start positions of the middle piece: [mid_lo, len - mid_hi]
};
struct __align__(16) IeGlobProbe {  // what the per-pattern probe needs, in two 16-byte constant loads
    uint32_t pre_word, pre_mask;   // the probe word of the prefix image (word `probe` of pre / pre_mask)
    uint32_t suf_word, suf_mask;   // the last word of the suffix image
    uint32_t min_len, max_len;     // len must lie in [min_len, max_len] (max_len = min_len: no star)
    uint32_t probe_off;            // byte offset of the prefix probe word in the key
    uint32_t kind;                 // IE_GLOB_* | complete << 8
};
struct IeGlobPatterns {  // passed by value as a kernel parameter (about 12 KiB; CUDA >= 12.1 allows 32 KiB)
    uint32_t n_pat;
    uint32_t invert;
    uint32_t any_pre, any_suf;
    uint16_t off[IE_MAX_PATTERNS + 1];
    uint8_t bytes[3584];
    uint8_t pad_[14];  // probe[] is 16-byte aligned
    IeGlobProbe probe[IE_MAX_PATTERNS];
    IeGlobFast fast[IE_MAX_PATTERNS];
};
void ie_glob_compile(IeGlobPatterns* pats);  // host: fills fast[] / any_pre / any_suf from bytes / off
// d_first (may be NULL): [n] index of the first matching pattern, 0xFFFFFFFF when none
// a few (long) texts, any number / length of patterns in global memory: one CTA per text
cudaError_t ie_launch_glob_first_long(const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n, const uint8_t* d_pats,
                                      const uint64_t* d_pat_offs, uint32_t n_pat, uint32_t* d_first, cudaStream_t stream);
cudaError_t ie_launch_glob(const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n, const IeGlobPatterns& pats,
                           uint32_t* d_mask, uint64_t* d_n_deleted, uint32_t* d_first, cudaStream_t stream);

// ---- device-side table build and in-place mutation (ie_table_build.cu) ------------------------------------------------
struct IeTableHeader {     // lives behind the view array of a table
    uint64_t arena_used;   // bytes of the arena handed out so far (bump allocator)
    uint32_t flags;        // bit 0: some value holds properly nested groups of its own (rescan rounds can do work)
    uint32_t error;        // 1 bad input arrays, 2 probe chain exhausted, 4 arena full, 8 slot array too full for an insert
};
struct IeBuildArgs {
    uint8_t* base;               // the table allocation
    uint64_t arena_off, arena_bytes;
    IeTableHeader* hdr;
    uint32_t* used;              // [n_states] non-empty slots per snapshot
    const uint64_t* state_offs;  // [n_states + 1] (device copies of the caller's arrays from here on)
    const uint64_t* slot_base;   // [n_states] first slot of each snapshot, in slots from `base`
    const uint32_t* slot_cap;    // [n_states] power of two
    uint32_t* slot_of;           // [items] scratch: the slot each item probed to
    const uint8_t* keys; const uint64_t* key_offs;
    const uint8_t* vals; const uint64_t* val_offs;
    const uint8_t* tags;
    const uint8_t* clock;        // "HH:MM" at 0, "HH:MM:SS" at 8, their renderings at 16 and 80 (<= 64 bytes each)
    uint32_t hhmm_len, hhmmss_len;
    uint64_t n;                  // caller inserts (all snapshots)
    uint32_t n_states;
    uint32_t with_clock;         // bit 0: "HH:MM" given, bit 1: "HH:MM:SS" given
};
struct IeMutateArgs {
    uint8_t* base;
    uint64_t arena_off, arena_bytes;
    IeTableHeader* hdr;
    const IeTableView* views;
    uint32_t* used_slots;        // [n_states] non-empty slots (live + tombstones)
    uint32_t state, all_states, n_ops;
    const uint8_t* keys; const uint64_t* key_offs;
    const uint8_t* vals; const uint64_t* val_offs;  // vals == nullptr: delete
    const uint8_t* tags; const uint8_t* flags;      // per op: JSON type, IE_VF_* of the value (classified by the host)
    const uint32_t* entries;                        // per op: what a typed result reports in aux (nullptr: IE_AUX_NONE)
};
cudaError_t ie_launch_table_build(const IeBuildArgs& a, IeTableView* d_views, cudaStream_t stream);
cudaError_t ie_launch_table_mutate(const IeMutateArgs& m, uint32_t n_states, cudaStream_t stream);
