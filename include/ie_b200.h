/* ie_b200.h — C ABI of the B200-native batched `{key}` interpolation engine.
 *
 * Drop-in boundary for the resolver of tillfalko/interpolation-engine.  The reference has no
 * FFI of its own; the boundary is the Rust module API of rust-project/src/interp.rs plus the
 * two wildcard helpers of rust-project/src/runtime.rs.  Every entry point below names the
 * reference item (file:line) it replaces.  All functions return an ie_status_t (0 = IE_OK) and
 * never throw or abort across the boundary; ie_last_error() returns the message for the calling
 * thread.  There is no CPU fallback: every resolve/escape/glob call runs CUDA kernels on the
 * engine's device and fails with IE_E_CUDA when no device is usable.
 *
 * Threading: an engine owns its streams, staging buffers and workspace; calls on ONE engine must not
 * overlap (the reference calls the resolver from one thread, interp.rs is synchronous).  Use one engine
 * per host thread or per GPU; tables belong to the engine that packed them.  Host-buffer results stay
 * owned by the engine until its next call.  Device-buffer calls are asynchronous on the given stream
 * and share the engine's workspace (counters, index lists, the general path's scratch): ONE device-buffer call
 * in flight per engine - issue them on one stream, or order them with events; a second engine on the same
 * device gives a second workspace.
 *
 * Data layout conventions
 *   string arenas   : `bytes` + `offs[n+1]` (uint64, offs[0]=0, offs[n]=total bytes), no separators
 *   packed inserts  : key arena + value arena + tags[n]; a value is the value_to_string()
 *                     rendering (interp.rs:314-322) of the serde_json::Value, the tag its type
 *   results         : out arena + out_offs[n] + out_lens[n] + status[n] + aux[n]
 */
#ifndef IE_B200_H
#define IE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- call status --------------------------------------------------------------------------- */
typedef enum {
    IE_OK = 0,
    IE_E_INVALID = 1,  /* bad argument */
    IE_E_CUDA = 2,     /* CUDA runtime error (no device, launch failure, ...) */
    IE_E_NOMEM = 3,    /* host or device allocation failed */
    IE_E_OVERFLOW = 4, /* caller-provided output arena too small; see ie_batch_info.out_bytes */
} ie_status_t;

/* ---- per-template result status (low 8 bits of status[i]; bits 8..15 = value tag) ----------- */
enum {
    IE_RES_STRING = 0,      /* Ok(Value::String(out bytes))                          interp.rs:88 */
    IE_RES_TYPED = 1,       /* simple path: Ok(inserts[key].clone()); out bytes = value_to_string,
                               aux = entry index of the insert (n, n+1 = "HH:MM", "HH:MM:SS")  :45-52 */
    IE_RES_UNEVEN = 2,      /* Err("Interpolation error: uneven number of '{' and '}' in: " + out) :58 */
    IE_RES_UNSUPPORTED = 3, /* Err("Trying to interpolate '" + out + "' of unsupported type")     :76 */
    IE_RES_EMPTY_KEY = 4,   /* Err("Tried to interpolate empty string ''")                        :105 */
    IE_RES_ARG_MISSING = 5, /* Err("Argument interpolation key '" + out + "' is used but not provided") :113 */
    IE_RES_NOT_FOUND = 6,   /* Err("Could not find variable '" + out + "'")                       :136 */
    IE_RES_PANIC = 7,       /* the reference panics: unwrap() on "no '}' after the last '{'"      :63-66 */
    IE_RES_LIMIT = 8,       /* expansion bound hit (the reference would loop forever)              */
};
#define IE_RES_CODE(s) ((int)((s) & 0xFF))
#define IE_RES_TAG(s) ((int)(((s) >> 8) & 0xFF))
#define IE_AUX_NONE 0xFFFFFFFFu

/* value type tags (serde_json::Value variants) */
enum { IE_TAG_NULL = 0, IE_TAG_BOOL = 1, IE_TAG_NUMBER = 2, IE_TAG_STRING = 3, IE_TAG_ARRAY = 4, IE_TAG_OBJECT = 5 };

typedef struct ie_engine ie_engine; /* one per (process, device): stream, staging, scratch      */
typedef struct ie_table ie_table;   /* device-resident packed `inserts` map (immutable snapshot) */

/* Expansion bounds.  The reference has none (a self-referential insert loops forever,
 * interp.rs:54); exceeding one yields IE_RES_LIMIT for that template only.
 * A field left 0 takes its default, and the defaults are only where a template STARTS: the host-buffer
 * calls (ie_resolve_batch and everything built on it) re-run a template that hit a default bound with
 * 8x the bound, again and again, up to IE_HARD_MAX_EXPANSIONS / IE_HARD_MAX_RESULT_BYTES, so that every
 * input on which the reference terminates within those caps resolves instead of reporting a limit.
 * A bound the caller sets explicitly is final.  The device-buffer call never escalates (no host in the
 * loop): ie_batch_info.n_limit tells the caller how many templates stopped at its bounds. */
#define IE_HARD_MAX_EXPANSIONS (1u << 17)
#define IE_HARD_MAX_RESULT_BYTES (1u << 26)
typedef struct {
    uint32_t max_expansions;   /* lookups per template on the general path (default 4096)       */
    uint32_t max_result_bytes; /* bytes of intermediate/final text per template on the general
                                  path (default 64 KiB)                                         */
    uint32_t avg_template_bytes; /* device-buffer calls only: mean template length, used to size the
                                  CTA tiles (0 = short templates, <= 230 bytes); the host-buffer
                                  calls measure it themselves                                   */
    uint32_t avg_template_groups; /* device-buffer calls only: mean number of {...} groups per template (0 = at
                                  most 4); with avg_template_bytes it sizes the CTA tiles so that dense templates
                                  stay on the fast path.  The host-buffer calls estimate it from the text.   */
    uint32_t rescan_rounds;    /* interp.rs:81-83 rescans every spliced value.  Values whose own groups nest
                                  properly are resolved by running the template through the fast kernel again
                                  ("round"), up to this many times (max 3); what is left, and every other kind
                                  of rescan, takes the general path.  0 = default: 2 for the host-buffer calls,
                                  none for the device-buffer call (fewer launches; same results either way) */
} ie_limits;

typedef struct {
    uint64_t n;           /* templates processed */
    uint64_t out_bytes;   /* bytes used in the out arena (needed size when IE_E_OVERFLOW)        */
    uint64_t n_general;   /* templates that took the general (slow) path                         */
    uint64_t n_limit;     /* templates whose status is IE_RES_LIMIT                               */
    float kernel_ms;      /* device time of the resolve kernels (CUDA events), host API only     */
} ie_batch_info;

const char* ie_last_error(void);
int ie_device_count(void);

/* ---- engine ------------------------------------------------------------------------------- */
ie_status_t ie_engine_create(int device, ie_engine** out);
void ie_engine_destroy(ie_engine* e);
void* ie_engine_stream(ie_engine* e); /* cudaStream_t the engine launches on */
ie_status_t ie_engine_sync(ie_engine* e);

/* ---- inserts snapshot -> device table ------------------------------------------------------
 * Replaces the `&Map<String, Value>` argument of interp.rs:31/:91/:179 (callers pass a snapshot
 * clone, runtime.rs:700).  Later duplicates of a key win (Map::insert semantics, interp.rs:139).
 * `hhmm` / `hhmmss` (may be NULL) are the renderings of the special keys "HH:MM" / "HH:MM:SS"
 * (interp.rs:96-104); when given they shadow inserts of the same name, as in the reference. */
ie_status_t ie_table_pack(ie_engine* e, uint64_t n, const uint8_t* keys, const uint64_t* key_offs,
                          const uint8_t* vals, const uint64_t* val_offs, const uint8_t* tags,
                          const char* hhmm, const char* hhmmss, ie_table** out);
/* Many snapshots at once (the cloned states of one program, runtime.rs:700 called once per state):
 * snapshot s owns the inserts [state_offs[s], state_offs[s+1]) of the packed arrays.  One device
 * allocation; the tables are built on the device (see ie_table_build_ms below).  A table that holds S
 * snapshots makes every ie_resolve_batch* call a cross product: all n templates are resolved against
 * every snapshot, result index = s * n + template, and every result array holds S * n entries.
 * `aux` entry indices are relative to the snapshot's first insert. */
ie_status_t ie_table_pack_many(ie_engine* e, uint64_t n_states, const uint64_t* state_offs, const uint8_t* keys,
                               const uint64_t* key_offs, const uint8_t* vals, const uint64_t* val_offs,
                               const uint8_t* tags, const char* hhmm, const char* hhmmss, ie_table** out);
/* Snapshots of 4096 inserts and more, and every ie_table_pack_many table, are built ON THE DEVICE: the packed arrays
 * are uploaded as they are and one thread per insert hashes its key, claims a slot, classifies and copies its value
 * (ie_table_build.cu); the host only lays out capacities.  ie_table_build_ms = device time of those kernels (0 for a
 * host-built table). */
double ie_table_build_ms(const ie_table* t);

/* ---- set_interpdata / delete_interpdata (interp.rs:139-145) on a packed table, in place --------------------------
 * The reference mutates its inserts map between tasks (17 set_interpdata call sites in runtime.rs) and snapshots it
 * per task (runtime.rs:700); with these two calls a packed table follows the map without being packed again.
 * `state` selects the snapshot of an ie_table_pack_many table (IE_ALL_STATES: the same operations on every snapshot).
 * Operations apply in order (a later one on the same key wins).  ie_table_set: insert or overwrite; `tags[i]` is the
 * JSON type of value i (its text is the value_to_string rendering, as in ie_table_pack), `entries[i]` (may be NULL)
 * what `aux` reports when a typed simple-path result is this insert.  ie_table_delete: keys that are absent are
 * ignored, like Map::remove.  Both are synchronous and must not overlap other work on the table.
 * IE_E_OVERFLOW: a snapshot's slot array is more than 3/4 full or the table's arena (which keeps 1/8 spare room,
 * at least 4 KiB) is exhausted; the operations that did not fit were skipped, everything else was applied and the
 * table is consistent - pack the snapshot again. */
#define IE_ALL_STATES 0xFFFFFFFFu
ie_status_t ie_table_set(ie_engine* e, ie_table* t, uint32_t state, uint64_t n, const uint8_t* keys, const uint64_t* key_offs,
                         const uint8_t* vals, const uint64_t* val_offs, const uint8_t* tags, const uint32_t* entries);
ie_status_t ie_table_delete(ie_engine* e, ie_table* t, uint32_t state, uint64_t n, const uint8_t* keys, const uint64_t* key_offs);

/* A table must outlive the work that reads it: the host-buffer calls are synchronous, but after a *_device call on a
 * caller's stream free the table only once that stream has passed the call (stream synchronise or an event): tables
 * up to 48 MiB live in the device's stream-ordered pool and are released in the order of the ENGINE's stream, which
 * does not see work queued elsewhere. */
void ie_table_free(ie_table* t);
uint64_t ie_table_device_bytes(const ie_table* t);
uint32_t ie_table_states(const ie_table* t); /* snapshots held (1 for ie_table_pack) */

/* ---- interpolate_inserts, batched (interp.rs:31-89) -----------------------------------------
 * Host-buffer form: copies the template arena to the device, resolves all n templates against
 * `t`, copies results back.  Result pointers stay owned by the engine and are valid until the
 * next call on the same engine. */
typedef struct {
    const uint8_t* out;       /* result bytes */
    const uint64_t* out_offs; /* [n] start of result i in `out` (tiles claim arena ranges in completion
                                 order: positions are not monotone in i, and gaps of < 16 bytes separate tiles) */
    const uint32_t* out_lens; /* [n] length of result i */
    const int32_t* status;    /* [n] IE_RES_* | tag << 8 */
    const uint32_t* aux;      /* [n] insert entry index for IE_RES_TYPED */
    ie_batch_info info;
} ie_result;
/* tmpl_offs[0] need not be 0: (tmpl, tmpl_offs + k) is a shard of a larger arena, only its bytes are copied. */
ie_status_t ie_resolve_batch(ie_engine* e, const ie_table* t, const uint8_t* tmpl, const uint64_t* tmpl_offs,
                             uint64_t n, const ie_limits* limits, ie_result* res);

/* Several GPUs of one box, one process (SURVEY.md §8 e): templates are independent given an immutable snapshot, so
 * ONE host batch is cut into contiguous shards of ceil(n / n_engines) templates, shard g goes through engines[g] (one
 * host thread and one set of streams per device; tables[g] = the snapshot packed on that engine) and there is no
 * device-to-device traffic.  shards[g] describes shard g: templates [first, first + n), results in res (owned by
 * engines[g] until its next call, like ie_resolve_batch's).  ie_shards_gather is the host gather: it concatenates the
 * shards' results in template order into caller arrays (out_offs gets n + 1 entries, contiguous). */
typedef struct {
    uint64_t first, n;   /* the shard's templates */
    ie_result res;       /* indices relative to `first` */
    ie_status_t status;  /* of this shard's call */
    char error[160];
} ie_shard_result;
ie_status_t ie_resolve_batch_multi(ie_engine* const* engines, const ie_table* const* tables, uint32_t n_engines,
                                   const uint8_t* tmpl, const uint64_t* tmpl_offs, uint64_t n, const ie_limits* limits,
                                   ie_shard_result* shards);
ie_status_t ie_shards_gather(const ie_shard_result* shards, uint32_t n_shards, uint8_t* out, uint64_t out_capacity,
                             uint64_t* out_offs, int32_t* status, uint32_t* aux, uint64_t* out_bytes);

/* Device-buffer form (inputs and outputs already resident in HBM; asynchronous on `stream`,
 * which may be NULL for the engine's own stream).  `d_info` is a device ie_batch_info. */
/* When the out arena is too small, d_info->out_bytes > out_capacity afterwards; it is then a LOWER bound of the
 * need (stages behind the overflowing one are skipped) and results whose range lies beyond out_capacity were not
 * written: regrow to at least twice the capacity and rerun, as ie_resolve_batch does. */
ie_status_t ie_resolve_batch_device(ie_engine* e, const ie_table* t, const uint8_t* d_tmpl, const uint64_t* d_tmpl_offs,
                                    uint64_t n, const ie_limits* limits, uint8_t* d_out, uint64_t out_capacity,
                                    uint64_t* d_out_offs, uint32_t* d_out_lens, int32_t* d_status, uint32_t* d_aux,
                                    ie_batch_info* d_info, void* stream);

/* ---- get_interpdata for a batch of literal keys (interp.rs:91-137; map probe only: the caller
 * handles "" and the ARG<digits> error text).  tag_out[i] = IE_TAG_* or -1 on a miss;
 * entry_out[i] = insert index (n, n+1 = the clock keys) or IE_AUX_NONE. */
ie_status_t ie_lookup_batch(ie_engine* e, const ie_table* t, const uint8_t* keys, const uint64_t* key_offs, uint64_t n,
                            int32_t* tag_out, uint32_t* entry_out);

/* ---- recursive_unescape / recursive_escape on strings, batched (interp.rs:147-177; also the
 * inline replaces of `print` / `write`, runtime.rs:1053-1055, 1272) --------------------------
 * mode 0: unescape ("\{" -> "{", then "\}" -> "}");  mode 1: escape ("{" -> "\{", "}" -> "\}"). */
ie_status_t ie_escape_batch(ie_engine* e, int mode, const uint8_t* in, const uint64_t* in_offs, uint64_t n,
                            const uint8_t** out, const uint64_t** out_offs);
/* device arenas: in_bytes >= d_in_offs[n] - d_in_offs[0] (the caller knows its arena; it sizes the
 * tile bookkeeping without a device->host read); d_out_offs gets n + 1 entries.  The call is asynchronous
 * and cannot report a too-small arena through its return value: d_out_offs[n] always receives the bytes
 * the output needs, and when it exceeds out_capacity the arena and the other offsets are incomplete
 * (escape at most doubles its input, unescape never grows it: size the arena accordingly). */
ie_status_t ie_escape_batch_device(ie_engine* e, int mode, const uint8_t* d_in, const uint64_t* d_in_offs, uint64_t n,
                                   uint64_t in_bytes, uint8_t* d_out, uint64_t out_capacity, uint64_t* d_out_offs,
                                   void* stream);

/* ---- wildcard_match over a key set: `delete` / `delete_except` (runtime.rs:1198-1239, 1633-1647)
 * mask bit k (uint32 words, LSB first) = key k is deleted, i.e. (any pattern matches) != invert.
 * Survivors keep their (sorted) input order.  n_pat <= IE_MAX_PATTERNS. */
#define IE_MAX_PATTERNS 64
ie_status_t ie_glob_sweep(ie_engine* e, const uint8_t* keys, const uint64_t* key_offs, uint64_t n,
                          const uint8_t* pats, const uint64_t* pat_offs, uint32_t n_pat, int invert,
                          uint32_t* mask, uint64_t* n_deleted);
ie_status_t ie_glob_sweep_device(ie_engine* e, const uint8_t* d_keys, const uint64_t* d_key_offs, uint64_t n,
                                 const uint8_t* pats, const uint64_t* pat_offs, uint32_t n_pat, int invert,
                                 uint32_t* d_mask, uint64_t* d_n_deleted, void* stream);

/* First-match form for goto_map / replace_map (runtime.rs:1085-1133, 1649-1692: the FIRST wildcard of an
 * ordered list that matches wins): first[k] = index of the first pattern matching key k, 0xFFFFFFFF if none.
 * Up to 256 keys (the callers above test ONE text, possibly kilobytes long) run one CTA per key with the patterns
 * in global memory: any number and length of patterns.  More keys take the sweep kernel and its limits
 * (IE_MAX_PATTERNS patterns, 3584 bytes of pattern text). */
ie_status_t ie_glob_first_match(ie_engine* e, const uint8_t* keys, const uint64_t* key_offs, uint64_t n,
                                const uint8_t* pats, const uint64_t* pat_offs, uint32_t n_pat, uint32_t* first);

/* ---- device memory helpers for hosts without a CUDA binding (tests, bench, FFI callers) ----- */
ie_status_t ie_device_alloc(ie_engine* e, uint64_t bytes, void** d_ptr);
void ie_device_free(ie_engine* e, void* d_ptr);
ie_status_t ie_copy_to_device(ie_engine* e, void* d_dst, const void* h_src, uint64_t bytes);
ie_status_t ie_copy_to_host(ie_engine* e, void* h_dst, const void* d_src, uint64_t bytes);
ie_status_t ie_host_alloc(uint64_t bytes, void** h_ptr); /* page-locked host memory for the host-buffer calls */
/* write-combined page-locked memory: for INPUT arenas the host only writes (reads from it are slow); the device reads
 * it without snooping the CPU caches */
ie_status_t ie_host_alloc_wc(uint64_t bytes, void** h_ptr);
void ie_host_free(void* h_ptr);

/* ---- JSON-level mirror of the interp.rs public functions (host logic above the batch ABI) ----
 * args: {"fn": name, "inserts": {...}, "content"|"key"|"value"|"pattern"|"text"|"wildcards": ...}
 * with fn one of interpolate_inserts (:31), get_simple_insertkey (:11), get_interpdata (:91),
 * recursive_interpolate (:179), recursive_escape (:163), recursive_unescape (:147),
 * extract_insert_keys (:248), value_to_string (:314), wildcard_match (runtime.rs:1633),
 * wildcard_captures (runtime.rs:1754), delete / delete_except (runtime.rs:1198 / 1219),
 * replace_map (runtime.rs:1649; args item, wildcard_maps, repeat_until_done) and goto_map (runtime.rs:1085-1133;
 * args text, target_maps -> {"value", "target", "interpolation_error"}), add_line_numbers / load_program
 * (parser.rs:74 / :8; arg text; host only), interpolation_trace (arg value: the strings recursive_interpolate sends
 * to the resolver, in order).
 * A caller that keeps ONE inserts map over many calls (the reference's run loop) registers it once -
 * {"fn": "snapshot_create", "inserts": {...}} -> id - mirrors its set_interpdata / delete_interpdata calls with
 * {"fn": "snapshot_set", "snapshot": id, "key", "value"} / {"fn": "snapshot_delete", "snapshot": id, "key"} (the
 * packed device table is patched in place, ie_table_set / ie_table_delete) and passes {"snapshot": id} instead of
 * "inserts" to every function above: no per-call serialisation, packing or upload of the state.  "snapshot_free"
 * releases it (ie_engine_destroy releases what is left).  Returns malloc'ed UTF-8 JSON
 * {"ok": value} | {"err": {"code", "message", "payload"}}; free with ie_free. */
ie_status_t ie_call_json(ie_engine* e, const char* args_json, size_t len, char** out_json, size_t* out_len);
void ie_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* IE_B200_H */
