// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (interpolation_engine_b200/) never links or loads it.
//
// C entry points over the restatement in oracle_interp.hpp:
//   * orc_call_json      — one JSON-in/JSON-out dispatcher mirroring every pub fn of
//                          rust-project/src/interp.rs plus wildcard_match / wildcard_captures /
//                          delete / delete_except (runtime.rs:1633, 1754, 1198, 1219);
//   * orc_resolve_batch  — the arena-level batch the CUDA path is compared with and the
//                          multithreaded CPU baseline bench.py times;
//   * orc_glob_sweep     — wildcard delete sweep over a key arena (runtime.rs:1198-1239).
#include <atomic>
#include <cstdlib>
#include <thread>

#include "oracle_interp.hpp"

using namespace orc;

namespace {

const Value* field(const Value& obj, const char* name) {
    if (obj.kind != Value::Obj) return nullptr;
    auto it = obj.o->find(name);
    return it == obj.o->end() ? nullptr : &it->second;
}
std::string str_field(const Value& obj, const char* name) {
    const Value* v = field(obj, name);
    if (!v || v->kind != Value::String) throw std::runtime_error(std::string("missing string field ") + name);
    return v->s;
}
Ctx ctx_from(const Value& args) {
    Ctx ctx;
    if (const Value* c = field(args, "clock")) {
        if (const Value* a = field(*c, "hhmm")) ctx.fixed_hhmm = a->s;
        if (const Value* b = field(*c, "hhmmss")) ctx.fixed_hhmmss = b->s;
    }
    if (const Value* d = field(args, "inserts_dir")) if (d->kind == Value::String) ctx.inserts_dir = d->s;
    if (const Value* m = field(args, "max_iterations")) ctx.max_iterations = (long)m->i;
    if (const Value* m = field(args, "max_bytes")) ctx.max_bytes = (size_t)m->i;
    return ctx;
}
char* dup_out(const std::string& s, size_t* out_len) {
    char* p = (char*)std::malloc(s.size() + 1);
    std::memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    if (out_len) *out_len = s.size();
    return p;
}
Value strings_value(const std::vector<std::string>& v) {
    Array a;
    for (auto& s : v) a.push_back(Value::string(s));
    return Value::array(std::move(a));
}

Value dispatch(const Value& args) {
    std::string fn = str_field(args, "fn");
    static const Object empty;
    const Value* insv = field(args, "inserts");
    const Object& inserts = (insv && insv->kind == Value::Obj) ? *insv->o : empty;
    Ctx ctx = ctx_from(args);
    if (fn == "interpolate_inserts") return interpolate_inserts(inserts, str_field(args, "content"), ctx);
    if (fn == "get_simple_insertkey") {
        auto k = get_simple_insertkey(str_field(args, "content"));
        return k ? Value::string(*k) : Value::null();
    }
    if (fn == "get_interpdata") return get_interpdata(inserts, str_field(args, "key"), ctx);
    if (fn == "recursive_interpolate") return recursive_interpolate(inserts, *field(args, "value"), ctx);
    if (fn == "recursive_escape") return recursive_escape(*field(args, "value"));
    if (fn == "recursive_unescape") return recursive_unescape(*field(args, "value"));
    if (fn == "value_to_string") return Value::string(value_to_string(*field(args, "value")));
    if (fn == "extract_insert_keys") { std::vector<std::string> k; extract_insert_keys(*field(args, "value"), k); return strings_value(k); }
    if (fn == "wildcard_match") return Value::boolean(wildcard_match(str_field(args, "pattern"), str_field(args, "text")));
    if (fn == "wildcard_captures") return strings_value(wildcard_captures(str_field(args, "pattern"), str_field(args, "text")));
    if (fn == "delete" || fn == "delete_except") {
        Object ins = inserts;
        const Value* w = field(args, "wildcards");
        std::vector<Value> wl;
        if (w && w->kind == Value::Arr) wl = *w->a;
        auto deleted = delete_matching(ins, wl, fn == "delete_except");
        Object out;
        out["deleted"] = strings_value(deleted);
        out["inserts"] = Value::object(std::move(ins));
        return Value::object(std::move(out));
    }
    if (fn == "interpolation_trace") {  // the resolver calls of recursive_interpolate(value), in order (interp.rs:179-246)
        std::vector<std::string> t, l;
        interpolation_trace(*field(args, "value"), t, l);
        Object out;
        out["templates"] = strings_value(t);
        out["lookups"] = strings_value(l);
        return Value::object(std::move(out));
    }
    if (fn == "add_line_numbers") return Value::string(add_line_numbers(str_field(args, "text")));  // parser.rs:74
    if (fn == "load_program") return load_program(str_field(args, "text"));                          // parser.rs:8
    if (fn == "replace_map") {  // runtime.rs:1649
        const Value* maps = field(args, "wildcard_maps");
        if (!maps || maps->kind != Value::Arr) throw task_error("replace_map.wildcard_maps must be array");
        const Value* item = field(args, "item");
        const Value* rep = field(args, "repeat_until_done");
        return replace_map(item ? *item : Value::null(), *maps->a, inserts, ctx, rep && rep->kind == Value::Bool && rep->b);
    }
    if (fn == "goto_map") {  // runtime.rs:1085
        const Value* maps = field(args, "target_maps");
        if (!maps || maps->kind != Value::Arr) throw task_error("goto_map.target_maps must be array");
        const GotoChoice g = goto_map(str_field(args, "text"), *maps->a, inserts, ctx);
        Object out;
        out["value"] = Value::string(g.value_text);
        out["target"] = Value::string(g.target);
        out["interpolation_error"] = Value::boolean(g.interp_error);
        return Value::object(std::move(out));
    }
    throw std::runtime_error("unknown fn " + fn);
}

}  // namespace

extern "C" {

void orc_free(void* p) { std::free(p); }

// args_json: {"fn": "...", ...}; returns malloc'ed JSON {"ok": v} | {"err": {code,message,payload}}
char* orc_call_json(const char* args_json, size_t len, size_t* out_len) {
    Object res;
    try {
        Value args = parse_json(std::string(args_json, len));
        res["ok"] = dispatch(args);
    } catch (const InterpError& e) {
        Object err;
        err["code"] = Value::integer(e.code);
        err["message"] = Value::string(e.message);
        err["payload"] = Value::string(e.payload);
        res["err"] = Value::object(std::move(err));
    } catch (const std::exception& e) {
        Object err;
        err["code"] = Value::integer(-1);
        err["message"] = Value::string(e.what());
        err["payload"] = Value::string("");
        res["err"] = Value::object(std::move(err));
    }
    return dup_out(to_json(Value::object(std::move(res))), out_len);
}

// Status codes shared with include/ie_b200.h: 0 string, 1 typed (simple path), 2.. = ErrCode.
// tags: 0 null, 1 bool, 2 number, 3 string, 4 array, 5 object; values are value_to_string text.
struct OrcTable { Object map; std::map<std::string, uint32_t> index; };

void* orc_table_build(uint64_t n, const uint8_t* keys, const uint64_t* key_offs,
                      const uint8_t* vals, const uint64_t* val_offs, const uint8_t* tags) {
    auto* t = new OrcTable();
    static const Value::Kind kinds[6] = {Value::Null, Value::Bool, Value::Int, Value::String, Value::Arr, Value::Obj};
    for (uint64_t i = 0; i < n; ++i) {
        std::string k((const char*)keys + key_offs[i], key_offs[i + 1] - key_offs[i]);
        std::string v((const char*)vals + val_offs[i], val_offs[i + 1] - val_offs[i]);
        Value val = tags[i] == 3 ? Value::string(std::move(v)) : Value::rendered(kinds[tags[i] % 6], std::move(v));
        t->map[k] = std::move(val);
        t->index[k] = (uint32_t)i;
    }
    return t;
}
void orc_table_free(void* t) { delete (OrcTable*)t; }

static int tag_of(const Value& v) {
    switch (v.kind) {
        case Value::Null: return 0; case Value::Bool: return 1; case Value::String: return 3;
        case Value::Arr: return 4; case Value::Obj: return 5; default: return 2;
    }
}

// Resolves n templates against one table with `threads` host threads (static contiguous
// partition).  out_offs has n+1 entries; *out_arena is malloc'ed (free with orc_free).
// status[i]: 0 string, 1 typed, else ErrCode; aux[i]: for typed results (tag<<28 | entry index
// when the key is a table entry, 0x0FFFFFFF for clock keys), else 0.
int orc_resolve_batch(void* table, const uint8_t* tmpl, const uint64_t* offs, uint64_t n, int threads,
                      const char* hhmm, const char* hhmmss,
                      uint8_t** out_arena, uint64_t* out_offs, int32_t* status, uint32_t* aux) {
    auto* t = (OrcTable*)table;
    Ctx ctx;
    if (hhmm) ctx.fixed_hhmm = hhmm;
    if (hhmmss) ctx.fixed_hhmmss = hhmmss;
    ctx.max_iterations = 4096;
    if (threads < 1) threads = 1;
    std::vector<std::vector<std::string>> outs((size_t)threads);
    auto work = [&](int tid) {
        uint64_t lo = n * (uint64_t)tid / (uint64_t)threads, hi = n * (uint64_t)(tid + 1) / (uint64_t)threads;
        auto& o = outs[(size_t)tid];
        o.reserve(hi - lo);
        for (uint64_t i = lo; i < hi; ++i) {
            std::string content((const char*)tmpl + offs[i], offs[i + 1] - offs[i]);
            aux[i] = 0;
            try {
                // A simple-path result keeps its type; recover which entry it was for aux.
                std::string sent = replace_all(replace_all(content, ESCAPED_START, REPLACED_START), ESCAPED_STOP, REPLACED_STOP);
                bool simple = get_simple_insertkey(sent).has_value();
                Value v = interpolate_inserts(t->map, content, ctx);
                if (simple) {
                    status[i] = 1;
                    aux[i] = ((uint32_t)tag_of(v) << 28) | 0x0FFFFFFFu;
                } else status[i] = 0;
                o.push_back(value_to_string(v));
            } catch (const InterpError& e) {
                status[i] = e.code;
                o.push_back(e.payload);
            }
        }
    };
    std::vector<std::thread> th;
    for (int k = 1; k < threads; ++k) th.emplace_back(work, k);
    work(0);
    for (auto& x : th) x.join();
    // the result arena is assembled by the same threads, each copying its own strings behind a prefix over the
    // per-thread totals (a serial 284 MB copy here was a third of the measured CPU time on 16 threads)
    std::vector<uint64_t> tbytes((size_t)threads + 1, 0), tfirst((size_t)threads + 1, 0);
    for (int k = 0; k < threads; ++k) {
        uint64_t b = 0;
        for (auto& s : outs[(size_t)k]) b += s.size();
        tbytes[(size_t)k + 1] = tbytes[(size_t)k] + b;
        tfirst[(size_t)k + 1] = tfirst[(size_t)k] + outs[(size_t)k].size();
    }
    const uint64_t total = tbytes[(size_t)threads];
    out_offs[n] = total;
    uint8_t* arena = (uint8_t*)std::malloc(total ? total : 1);
    auto copy = [&](int tid) {
        uint64_t at = tbytes[(size_t)tid], i = tfirst[(size_t)tid];
        for (auto& s : outs[(size_t)tid]) { out_offs[i++] = at; std::memcpy(arena + at, s.data(), s.size()); at += s.size(); }
    };
    std::vector<std::thread> tc;
    for (int k = 1; k < threads; ++k) tc.emplace_back(copy, k);
    copy(0);
    for (auto& x : tc) x.join();
    *out_arena = arena;
    return 0;
}

// bit k of mask (u32 words, little-endian bit order) = key k is deleted, i.e.
// (any pattern matches key k) != invert.   runtime.rs:1204, 1225.
int orc_glob_sweep(const uint8_t* keys, const uint64_t* key_offs, uint64_t n,
                   const uint8_t* pats, const uint64_t* pat_offs, uint32_t n_pat, int invert,
                   int threads, uint32_t* mask) {
    std::vector<std::string> patterns;
    for (uint32_t p = 0; p < n_pat; ++p) patterns.emplace_back((const char*)pats + pat_offs[p], pat_offs[p + 1] - pat_offs[p]);
    uint64_t words = (n + 31) / 32;
    if (threads < 1) threads = 1;
    auto work = [&](int tid) {
        uint64_t lo = words * (uint64_t)tid / (uint64_t)threads, hi = words * (uint64_t)(tid + 1) / (uint64_t)threads;
        for (uint64_t w = lo; w < hi; ++w) {
            uint32_t bits = 0;
            for (uint64_t k = w * 32; k < std::min<uint64_t>(n, w * 32 + 32); ++k) {
                std::string key((const char*)keys + key_offs[k], key_offs[k + 1] - key_offs[k]);
                bool any = false;
                for (auto& p : patterns) if (wildcard_match(p, key)) { any = true; break; }
                if (any != (invert != 0)) bits |= 1u << (k & 31);
            }
            mask[w] = bits;
        }
    };
    std::vector<std::thread> th;
    for (int k = 1; k < threads; ++k) th.emplace_back(work, k);
    work(0);
    for (auto& x : th) x.join();
    return 0;
}

// Test helper: string i of arena A lives at a[a_offs[i] .. + lens[i]), of arena B at b[b_offs[i] .. + lens[i]).
// Returns the first index whose bytes differ, or n when all are equal.
uint64_t orc_compare_ragged(const uint8_t* a, const uint64_t* a_offs, const uint8_t* b, const uint64_t* b_offs,
                            const uint32_t* lens, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i)
        if (lens[i] && std::memcmp(a + a_offs[i], b + b_offs[i], lens[i]) != 0) return i;
    return n;
}

}  // extern "C"
