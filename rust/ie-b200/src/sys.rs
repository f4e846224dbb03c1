//! Raw bindings: one declaration per entry point of `include/ie_b200.h` (same order).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub enum ie_engine {}
pub enum ie_table {}

pub const IE_OK: c_int = 0;

pub const IE_RES_STRING: i32 = 0;
pub const IE_RES_TYPED: i32 = 1;
pub const IE_RES_UNEVEN: i32 = 2;
pub const IE_RES_UNSUPPORTED: i32 = 3;
pub const IE_RES_EMPTY_KEY: i32 = 4;
pub const IE_RES_ARG_MISSING: i32 = 5;
pub const IE_RES_NOT_FOUND: i32 = 6;
pub const IE_RES_PANIC: i32 = 7;
pub const IE_RES_LIMIT: i32 = 8;

pub const IE_TAG_NULL: u8 = 0;
pub const IE_TAG_BOOL: u8 = 1;
pub const IE_TAG_NUMBER: u8 = 2;
pub const IE_TAG_STRING: u8 = 3;
pub const IE_TAG_ARRAY: u8 = 4;
pub const IE_TAG_OBJECT: u8 = 5;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct ie_limits {
    pub max_expansions: u32,
    pub max_result_bytes: u32,
    pub avg_template_bytes: u32,
    pub avg_template_groups: u32,
    pub rescan_rounds: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ie_batch_info {
    pub n: u64,
    pub out_bytes: u64,
    pub n_general: u64,
    pub n_limit: u64,
    pub kernel_ms: f32,
}

#[repr(C)]
pub struct ie_result {
    pub out: *const u8,
    pub out_offs: *const u64,
    pub out_lens: *const u32,
    pub status: *const i32,
    pub aux: *const u32,
    pub info: ie_batch_info,
}

pub type ie_status_t = c_int;

#[repr(C)]
pub struct ie_shard_result {
    pub first: u64,
    pub n: u64,
    pub res: ie_result,
    pub status: ie_status_t,
    pub error: [c_char; 160],
}

extern "C" {
    pub fn ie_last_error() -> *const c_char;
    pub fn ie_device_count() -> c_int;
    pub fn ie_engine_create(device: c_int, out: *mut *mut ie_engine) -> c_int;
    pub fn ie_engine_destroy(e: *mut ie_engine);
    pub fn ie_engine_stream(e: *mut ie_engine) -> *mut c_void;
    pub fn ie_engine_sync(e: *mut ie_engine) -> c_int;

    pub fn ie_table_pack(e: *mut ie_engine, n: u64, keys: *const u8, key_offs: *const u64, vals: *const u8, val_offs: *const u64,
                         tags: *const u8, hhmm: *const c_char, hhmmss: *const c_char, out: *mut *mut ie_table) -> c_int;
    pub fn ie_table_pack_many(e: *mut ie_engine, n_states: u64, state_offs: *const u64, keys: *const u8, key_offs: *const u64,
                              vals: *const u8, val_offs: *const u64, tags: *const u8, hhmm: *const c_char, hhmmss: *const c_char,
                              out: *mut *mut ie_table) -> c_int;
    pub fn ie_table_free(t: *mut ie_table);
    pub fn ie_table_device_bytes(t: *const ie_table) -> u64;
    pub fn ie_table_states(t: *const ie_table) -> u32;
    pub fn ie_table_build_ms(t: *const ie_table) -> f64;
    pub fn ie_table_set(e: *mut ie_engine, t: *mut ie_table, state: u32, n: u64, keys: *const u8, key_offs: *const u64,
                        vals: *const u8, val_offs: *const u64, tags: *const u8, entries: *const u32) -> ie_status_t;
    pub fn ie_table_delete(e: *mut ie_engine, t: *mut ie_table, state: u32, n: u64, keys: *const u8, key_offs: *const u64) -> ie_status_t;

    pub fn ie_resolve_batch(e: *mut ie_engine, t: *const ie_table, tmpl: *const u8, tmpl_offs: *const u64, n: u64,
                            limits: *const ie_limits, res: *mut ie_result) -> c_int;
    pub fn ie_resolve_batch_multi(engines: *const *mut ie_engine, tables: *const *const ie_table, n_engines: u32, tmpl: *const u8,
                                  tmpl_offs: *const u64, n: u64, limits: *const ie_limits, shards: *mut ie_shard_result) -> ie_status_t;
    pub fn ie_shards_gather(shards: *const ie_shard_result, n_shards: u32, out: *mut u8, out_capacity: u64, out_offs: *mut u64,
                            status: *mut i32, aux: *mut u32, out_bytes: *mut u64) -> ie_status_t;
    pub fn ie_resolve_batch_device(e: *mut ie_engine, t: *const ie_table, d_tmpl: *const u8, d_tmpl_offs: *const u64, n: u64,
                                   limits: *const ie_limits, d_out: *mut u8, out_capacity: u64, d_out_offs: *mut u64,
                                   d_out_lens: *mut u32, d_status: *mut i32, d_aux: *mut u32, d_info: *mut ie_batch_info,
                                   stream: *mut c_void) -> c_int;
    pub fn ie_lookup_batch(e: *mut ie_engine, t: *const ie_table, keys: *const u8, key_offs: *const u64, n: u64,
                           tag_out: *mut i32, entry_out: *mut u32) -> c_int;

    pub fn ie_escape_batch(e: *mut ie_engine, mode: c_int, input: *const u8, in_offs: *const u64, n: u64,
                           out: *mut *const u8, out_offs: *mut *const u64) -> c_int;
    pub fn ie_escape_batch_device(e: *mut ie_engine, mode: c_int, d_in: *const u8, d_in_offs: *const u64, n: u64, in_bytes: u64,
                                  d_out: *mut u8, out_capacity: u64, d_out_offs: *mut u64, stream: *mut c_void) -> c_int;

    pub fn ie_glob_sweep(e: *mut ie_engine, keys: *const u8, key_offs: *const u64, n: u64, pats: *const u8, pat_offs: *const u64,
                         n_pat: u32, invert: c_int, mask: *mut u32, n_deleted: *mut u64) -> c_int;
    pub fn ie_glob_sweep_device(e: *mut ie_engine, d_keys: *const u8, d_key_offs: *const u64, n: u64, pats: *const u8,
                                pat_offs: *const u64, n_pat: u32, invert: c_int, d_mask: *mut u32, d_n_deleted: *mut u64,
                                stream: *mut c_void) -> c_int;
    /// first[k] = index of the first pattern matching key k, 0xFFFF_FFFF if none (runtime.rs:1085-1133, 1649-1692)
    pub fn ie_glob_first_match(e: *mut ie_engine, keys: *const u8, key_offs: *const u64, n: u64, pats: *const u8, pat_offs: *const u64,
                               n_pat: u32, first: *mut u32) -> c_int;

    pub fn ie_device_alloc(e: *mut ie_engine, bytes: u64, d_ptr: *mut *mut c_void) -> c_int;
    pub fn ie_device_free(e: *mut ie_engine, d_ptr: *mut c_void);
    pub fn ie_copy_to_device(e: *mut ie_engine, d_dst: *mut c_void, h_src: *const c_void, bytes: u64) -> c_int;
    pub fn ie_copy_to_host(e: *mut ie_engine, h_dst: *mut c_void, d_src: *const c_void, bytes: u64) -> c_int;
    pub fn ie_host_alloc(bytes: u64, h_ptr: *mut *mut c_void) -> c_int;
    pub fn ie_host_alloc_wc(bytes: u64, h_ptr: *mut *mut c_void) -> ie_status_t;
    pub fn ie_host_free(h_ptr: *mut c_void);

    pub fn ie_call_json(e: *mut ie_engine, args_json: *const c_char, len: usize, out_json: *mut *mut c_char, out_len: *mut usize) -> c_int;
    pub fn ie_free(p: *mut c_void);
}
