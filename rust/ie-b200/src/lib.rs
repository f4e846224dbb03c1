//! Drop-in for `rust-project/src/interp.rs` of tillfalko/interpolation-engine.
//!
//! Same public items (names, argument meaning, `anyhow` error texts) as `interp.rs:7-322`; the work is done
//! by the B200 engine (`libie_b200.so`, `include/ie_b200.h`).  There is no CPU fallback: creating the engine
//! fails without a CUDA device.  Batch entry points (`Snapshot`, `interpolate_many`) are what a caller with many
//! templates or many cloned states should use; `LiveInserts` is the run loop's mutable map with its device table
//! patched in place; the single-call functions keep the `interp.rs` signatures so that `runtime.rs`, `math.rs`
//! and `analyzer.rs` compile unchanged.
//!
//! NOT COMPILED in the build container of this repository (no Rust toolchain there); the C ABI it binds is
//! covered by the Python parity tests in `tests/`.

pub mod sys;

use anyhow::{anyhow, Result};
use serde_json::{json, Map, Value};
use std::ffi::{CStr, CString};
use std::path::PathBuf;
use std::ptr;
use std::sync::OnceLock;

pub const INSERT_START: char = '{'; // interp.rs:7
pub const INSERT_STOP: char = '}'; // interp.rs:8
pub const ESCAPE: char = '\\'; // interp.rs:9

/// `model.rs:17-22`: only `inserts_dir` matters to the resolver.
#[derive(Clone, Debug, Default)]
pub struct ProgramLoadContext {
    pub inserts_dir: Option<PathBuf>,
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::ie_last_error()).to_string_lossy().into_owned() }
}

fn check(rc: i32) -> Result<()> {
    if rc == sys::IE_OK {
        Ok(())
    } else {
        Err(anyhow!("ie_b200: {}", last_error()))
    }
}

/// One engine per process and device (the reference is single threaded, `?Send` tasks on one thread).
pub struct Engine {
    raw: *mut sys::ie_engine,
}
unsafe impl Send for Engine {}
unsafe impl Sync for Engine {}

impl Engine {
    pub fn new(device: i32) -> Result<Engine> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::ie_engine_create(device, &mut raw) })?;
        Ok(Engine { raw })
    }

    /// JSON-level mirror (`ie_call_json`): `{"fn": name, ...}` -> `Ok(value)` or the reference's error text.
    fn call(&self, args: Value) -> Result<Value> {
        let text = CString::new(serde_json::to_string(&args)?)?;
        let (mut out, mut len) = (ptr::null_mut(), 0usize);
        check(unsafe { sys::ie_call_json(self.raw, text.as_ptr(), text.as_bytes().len(), &mut out, &mut len) })?;
        let bytes = unsafe { std::slice::from_raw_parts(out as *const u8, len) }.to_vec();
        unsafe { sys::ie_free(out as *mut _) };
        let mut reply: Value = serde_json::from_slice(&bytes)?;
        if let Some(ok) = reply.get_mut("ok") {
            return Ok(ok.take());
        }
        let msg = reply["err"]["message"].as_str().unwrap_or("ie_b200: malformed reply").to_string();
        if reply["err"]["code"].as_i64() == Some(sys::IE_RES_PANIC as i64) {
            panic!("{msg}"); // interp.rs:66 unwraps a None here
        }
        Err(anyhow!(msg))
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::ie_engine_destroy(self.raw) }
    }
}

/// The process-wide engine the single-call functions use (device 0, or `IE_B200_DEVICE`).
pub fn engine() -> &'static Engine {
    static ENGINE: OnceLock<Engine> = OnceLock::new();
    ENGINE.get_or_init(|| {
        let dev = std::env::var("IE_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        Engine::new(dev).expect("ie_b200: no usable CUDA device (this engine has no CPU fallback)")
    })
}

/// String arena in the layout of the C ABI: bytes + n+1 offsets.
#[derive(Default)]
struct Arena {
    bytes: Vec<u8>,
    offs: Vec<u64>,
}
impl Arena {
    fn new() -> Arena {
        Arena { bytes: Vec::new(), offs: vec![0] }
    }
    fn push(&mut self, s: &[u8]) {
        self.bytes.extend_from_slice(s);
        self.offs.push(self.bytes.len() as u64);
    }
    fn n(&self) -> u64 {
        (self.offs.len() - 1) as u64
    }
}

fn tag_of(v: &Value) -> u8 {
    match v {
        Value::Null => sys::IE_TAG_NULL,
        Value::Bool(_) => sys::IE_TAG_BOOL,
        Value::Number(_) => sys::IE_TAG_NUMBER,
        Value::String(_) => sys::IE_TAG_STRING,
        Value::Array(_) => sys::IE_TAG_ARRAY,
        Value::Object(_) => sys::IE_TAG_OBJECT,
    }
}

/// `interp.rs:314-322`, through the host mirror like every other function (one rendering of `serde_json`'s
/// number and object text: the library's).
pub fn value_to_string(value: &Value) -> String {
    match engine().call(json!({"fn": "value_to_string", "value": value})) {
        Ok(Value::String(s)) => s,
        _ => String::new(),
    }
}

/// An immutable inserts snapshot packed into the device hash table (`runtime.rs:700` clones the map once per
/// task; pack once per task and resolve every string of the task against it).
pub struct Snapshot {
    table: *mut sys::ie_table,
    values: Vec<Value>, // entry index -> original value (typed simple-path results, interp.rs:45-52)
    n_states: usize,
    per_state: Vec<usize>, // first entry of every state in `values`
}
unsafe impl Send for Snapshot {}

impl Snapshot {
    pub fn pack(engine: &Engine, inserts: &Map<String, Value>) -> Result<Snapshot> {
        Snapshot::pack_many(engine, std::slice::from_ref(inserts))
    }

    /// Many cloned states in one device allocation; resolving against it is a cross product
    /// (result index = state * n_templates + template).
    pub fn pack_many(engine: &Engine, states: &[Map<String, Value>]) -> Result<Snapshot> {
        let (mut keys, mut vals) = (Arena::new(), Arena::new());
        let mut tags = Vec::new();
        let mut values = Vec::new();
        let mut state_offs = vec![0u64];
        let mut per_state = Vec::new();
        for inserts in states {
            per_state.push(values.len());
            for (k, v) in inserts {
                // serde_json::Map without `preserve_order` is a BTreeMap: sorted key order
                keys.push(k.as_bytes());
                vals.push(value_to_string(v).as_bytes());
                tags.push(tag_of(v));
                values.push(v.clone());
            }
            state_offs.push(values.len() as u64);
        }
        // interp.rs:96-104: the clock keys are rendered once per call and shadow inserts of the same name
        let now = chrono::Local::now();
        let hhmm = CString::new(now.format("%H:%M").to_string())?;
        let hhmmss = CString::new(now.format("%H:%M:%S").to_string())?;
        let mut table = ptr::null_mut();
        check(unsafe {
            sys::ie_table_pack_many(engine.raw, states.len() as u64, state_offs.as_ptr(), keys.bytes.as_ptr(), keys.offs.as_ptr(),
                                    vals.bytes.as_ptr(), vals.offs.as_ptr(), tags.as_ptr(), hhmm.as_ptr(), hhmmss.as_ptr(), &mut table)
        })?;
        Ok(Snapshot { table, values, n_states: states.len(), per_state })
    }
}

impl Drop for Snapshot {
    fn drop(&mut self) {
        unsafe { sys::ie_table_free(self.table) }
    }
}

/// `interpolate_inserts` (interp.rs:31-89) for many templates in ONE launch.  With a snapshot of S states the
/// result vector holds S * contents.len() entries (state-major).  The inserts-dir fallback (interp.rs:122-134)
/// is not applied here; `interpolate_inserts` below routes through the JSON mirror, which does.
pub fn interpolate_many(engine: &Engine, snap: &Snapshot, contents: &[&str]) -> Result<Vec<Result<Value>>> {
    let mut arena = Arena::new();
    for c in contents {
        arena.push(c.as_bytes());
    }
    let mut res: sys::ie_result = unsafe { std::mem::zeroed() };
    check(unsafe { sys::ie_resolve_batch(engine.raw, snap.table, arena.bytes.as_ptr(), arena.offs.as_ptr(), arena.n(), ptr::null(), &mut res) })?;
    let total = contents.len() * snap.n_states;
    let mut out = Vec::with_capacity(total);
    for i in 0..total {
        let (off, len, status, aux) = unsafe {
            (*res.out_offs.add(i) as usize, *res.out_lens.add(i) as usize, *res.status.add(i), *res.aux.add(i) as usize)
        };
        let text = String::from_utf8_lossy(unsafe { std::slice::from_raw_parts(res.out.add(off), len) }).into_owned();
        let state = i / contents.len().max(1);
        out.push(match status & 0xFF {
            sys::IE_RES_STRING => Ok(Value::String(text)),
            sys::IE_RES_TYPED => match snap.values.get(snap.per_state[state] + aux) {
                Some(v) => Ok(v.clone()),
                None => Ok(Value::String(text)), // a clock key: rendered text
            },
            sys::IE_RES_UNEVEN => Err(anyhow!("Interpolation error: uneven number of '{{' and '}}' in: {text}")),
            sys::IE_RES_UNSUPPORTED => Err(anyhow!("Trying to interpolate '{text}' of unsupported type")),
            sys::IE_RES_EMPTY_KEY => Err(anyhow!("Tried to interpolate empty string ''")),
            sys::IE_RES_ARG_MISSING => Err(anyhow!("Argument interpolation key '{text}' is used but not provided")),
            sys::IE_RES_NOT_FOUND => Err(anyhow!("Could not find variable '{text}'")),
            sys::IE_RES_PANIC => panic!("called `Option::unwrap()` on a `None` value"),
            _ => Err(anyhow!("interpolation did not terminate within the engine's expansion limit")),
        });
    }
    Ok(out)
}

fn ctx_json(ctx: &ProgramLoadContext) -> Value {
    match &ctx.inserts_dir {
        Some(p) => json!(p.to_string_lossy()),
        None => Value::Null,
    }
}

/// interp.rs:11-29.
pub fn get_simple_insertkey(content: &str) -> Option<String> {
    match engine().call(json!({"fn": "get_simple_insertkey", "content": content})) {
        Ok(Value::String(s)) => Some(s),
        _ => None,
    }
}

/// The reference's run loop keeps ONE inserts map, mutates it between tasks (`set_interpdata`, 17 call sites in
/// `runtime.rs`) and resolves against it task after task (`runtime.rs:700`).  `LiveInserts` is that map with its
/// packed device table kept alive next to it: `set` / `delete` patch the table in place (`ie_table_set` /
/// `ie_table_delete` behind the mirror's "snapshot_set" / "snapshot_delete"), and every resolver call passes the
/// snapshot id instead of serialising the map.
pub struct LiveInserts {
    pub map: Map<String, Value>,
    id: u64,
}

impl LiveInserts {
    pub fn new(map: Map<String, Value>) -> Result<LiveInserts> {
        let id = engine().call(json!({"fn": "snapshot_create", "inserts": map}))?.as_u64().ok_or_else(|| anyhow!("ie_b200: malformed snapshot id"))?;
        Ok(LiveInserts { map, id })
    }
    /// interp.rs:139.
    pub fn set(&mut self, key: &str, value: Value) -> Result<()> {
        engine().call(json!({"fn": "snapshot_set", "snapshot": self.id, "key": key, "value": value}))?;
        self.map.insert(key.to_string(), value);
        Ok(())
    }
    /// interp.rs:143.
    pub fn delete(&mut self, key: &str) -> Result<()> {
        engine().call(json!({"fn": "snapshot_delete", "snapshot": self.id, "key": key}))?;
        self.map.remove(key);
        Ok(())
    }
    /// interp.rs:31.
    pub fn interpolate_inserts(&self, content: &str, ctx: &ProgramLoadContext) -> Result<Value> {
        engine().call(json!({"fn": "interpolate_inserts", "snapshot": self.id, "content": content, "inserts_dir": ctx_json(ctx)}))
    }
    /// interp.rs:179.
    pub fn recursive_interpolate(&self, value: Value, ctx: &ProgramLoadContext) -> Result<Value> {
        engine().call(json!({"fn": "recursive_interpolate", "snapshot": self.id, "value": value, "inserts_dir": ctx_json(ctx)}))
    }
    /// interp.rs:91.
    pub fn get_interpdata(&self, insertkey: &str, ctx: &ProgramLoadContext) -> Result<Value> {
        engine().call(json!({"fn": "get_interpdata", "snapshot": self.id, "key": insertkey, "inserts_dir": ctx_json(ctx)}))
    }
    /// runtime.rs:1649.
    pub fn replace_map(&self, item: Value, maps: &[Value], ctx: &ProgramLoadContext, repeat_until_done: bool) -> Result<Value> {
        engine().call(json!({"fn": "replace_map", "snapshot": self.id, "item": item, "wildcard_maps": maps,
                             "repeat_until_done": repeat_until_done, "inserts_dir": ctx_json(ctx)}))
    }
    /// runtime.rs:1085-1133.
    pub fn goto_map_target(&self, text: &str, target_maps: &[Value], ctx: &ProgramLoadContext) -> Result<String> {
        let reply = engine().call(json!({"fn": "goto_map", "snapshot": self.id, "text": text, "target_maps": target_maps, "inserts_dir": ctx_json(ctx)}))?;
        reply["target"].as_str().map(str::to_string).ok_or_else(|| anyhow!("ie_b200: malformed goto_map reply"))
    }
}

impl Drop for LiveInserts {
    fn drop(&mut self) {
        let _ = engine().call(json!({"fn": "snapshot_free", "snapshot": self.id}));
    }
}

/// interp.rs:31.
pub fn interpolate_inserts(inserts: &Map<String, Value>, content: &str, ctx: &ProgramLoadContext) -> Result<Value> {
    engine().call(json!({"fn": "interpolate_inserts", "inserts": inserts, "content": content, "inserts_dir": ctx_json(ctx)}))
}

/// interp.rs:91.
pub fn get_interpdata(inserts: &Map<String, Value>, insertkey: &str, ctx: &ProgramLoadContext) -> Result<Value> {
    engine().call(json!({"fn": "get_interpdata", "inserts": inserts, "key": insertkey, "inserts_dir": ctx_json(ctx)}))
}

/// interp.rs:139.
pub fn set_interpdata(inserts: &mut Map<String, Value>, key: &str, value: Value) {
    inserts.insert(key.to_string(), value);
}

/// interp.rs:143.
pub fn delete_interpdata(inserts: &mut Map<String, Value>, key: &str) {
    inserts.remove(key);
}

/// interp.rs:147 — strings and object keys, one escape-kernel batch per call.
pub fn recursive_unescape(value: Value) -> Value {
    engine().call(json!({"fn": "recursive_unescape", "value": value})).expect("ie_b200: recursive_unescape")
}

/// interp.rs:163.
pub fn recursive_escape(value: Value) -> Value {
    engine().call(json!({"fn": "recursive_escape", "value": value})).expect("ie_b200: recursive_escape")
}

/// interp.rs:179 — every string of the task goes into ONE resolve batch.
pub fn recursive_interpolate(inserts: &Map<String, Value>, value: Value, ctx: &ProgramLoadContext) -> Result<Value> {
    engine().call(json!({"fn": "recursive_interpolate", "inserts": inserts, "value": value, "inserts_dir": ctx_json(ctx)}))
}

/// interp.rs:248.
pub fn extract_insert_keys(value: &Value) -> Vec<String> {
    match engine().call(json!({"fn": "extract_insert_keys", "value": value})) {
        Ok(Value::Array(items)) => items.into_iter().filter_map(|v| v.as_str().map(str::to_string)).collect(),
        _ => Vec::new(),
    }
}

/// runtime.rs:1633 (private there; re-exported for `delete` / `delete_except` / `goto_map` / `replace_map`).
pub fn wildcard_match(pattern: &str, s: &str) -> bool {
    matches!(engine().call(json!({"fn": "wildcard_match", "pattern": pattern, "text": s})), Ok(Value::Bool(true)))
}

/// runtime.rs:1754 — the text each `*` of `pattern` swallows (greedy, leftmost), empty when there is no match.
pub fn wildcard_captures(pattern: &str, text: &str) -> Vec<String> {
    match engine().call(json!({"fn": "wildcard_captures", "pattern": pattern, "text": text})) {
        Ok(Value::Array(items)) => items.into_iter().filter_map(|v| v.as_str().map(str::to_string)).collect(),
        _ => Vec::new(),
    }
}

/// runtime.rs:1198-1239: keys of `inserts` to delete (`except` = delete_except), in sorted key order; one glob sweep.
pub fn delete_sweep(inserts: &Map<String, Value>, wildcards: &[String], except: bool) -> Result<Vec<String>> {
    let name = if except { "delete_except" } else { "delete" };
    let reply = engine().call(json!({"fn": name, "inserts": inserts, "wildcards": wildcards}))?;
    Ok(reply["deleted"].as_array().map(|a| a.iter().filter_map(|v| v.as_str().map(str::to_string)).collect()).unwrap_or_default())
}

/// runtime.rs:1649 — the `replace_map` task body.
pub fn replace_map(item: Value, maps: &[Value], inserts: &Map<String, Value>, ctx: &ProgramLoadContext, repeat_until_done: bool) -> Result<Value> {
    engine().call(json!({"fn": "replace_map", "inserts": inserts, "item": item, "wildcard_maps": maps,
                         "repeat_until_done": repeat_until_done, "inserts_dir": ctx_json(ctx)}))
}

/// runtime.rs:1085-1133 — the target selection of the `goto_map` task (the jump itself stays in the scheduler).
pub fn goto_map_target(text: &str, target_maps: &[Value], inserts: &Map<String, Value>, ctx: &ProgramLoadContext) -> Result<String> {
    let reply = engine().call(json!({"fn": "goto_map", "inserts": inserts, "text": text, "target_maps": target_maps, "inserts_dir": ctx_json(ctx)}))?;
    reply["target"].as_str().map(str::to_string).ok_or_else(|| anyhow!("ie_b200: malformed goto_map reply"))
}
