// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this code.  The product (interpolation_engine_b200/) never includes it.
//
// CPU restatement of the reference resolver, rust-project/src/interp.rs (all 322 lines),
// and of wildcard_match / wildcard_captures / delete / delete_except from
// rust-project/src/runtime.rs:1633-1647, 1754-1775, 1198-1239, and of the two callers that loop over the
// resolver, replace_map (runtime.rs:1649-1752) and goto_map's target selection (:1085-1133).  The pass structure of
// interpolate_inserts is kept on purpose (sentinel replace passes, per-iteration brace counts,
// rfind/find, whole-string rebuild per insertion, ordered-map lookup plus a deep value clone
// per hit) because this code is also the "reference-shaped" CPU baseline that bench.py times.
//
// PARITY STATUS: the Rust crate cannot be compiled in this environment (no cargo/rustc, no
// Cargo.lock, src/audio_web.rs missing) and the reference ships no tests or golden vectors.
// The restatement is pinned against (1) the reference's own Python twin executed here
// (oracle/gen_golden.py -> tests/golden/python_twin.json) on the subset where Python == Rust
// (SURVEY.md Appendix C), and (2) the hand-derived vectors of SURVEY.md Appendix B.
// Behaviour that neither source pins (f64 rendering, Object/Null stringification through
// serde_json, regex size limits) is PARITY UNPINNED.
#pragma once
#include <chrono>
#include <ctime>
#include <fstream>
#include <functional>
#include <optional>
#include <set>
#include <sstream>
#include <sys/stat.h>

#include "oracle_value.hpp"

namespace orc {

// interp.rs:7-9
constexpr char INSERT_START = '{';
constexpr char INSERT_STOP = '}';
constexpr char ESCAPE = '\\';

enum ErrCode {
    ERR_UNEVEN = 2,       // interp.rs:58-60
    ERR_UNSUPPORTED = 3,  // interp.rs:76-78
    ERR_EMPTY = 4,        // interp.rs:105
    ERR_ARG = 5,          // interp.rs:113-115
    ERR_NOT_FOUND = 6,    // interp.rs:136
    ERR_PANIC = 7,        // interp.rs:63-66 (unwrap on None)
    ERR_LIMIT = 8,        // not in the reference: it loops forever on self-referential values
    ERR_IO = 9,           // fs / json5 errors from the inserts-dir fallback (interp.rs:125-132)
};

struct InterpError {
    int code;
    std::string payload;  // the key, or the offending string for ERR_UNEVEN
    std::string message;  // exact anyhow message text
};

// model.rs:17-22 (only inserts_dir reaches the resolver) + test hooks
struct Ctx {
    std::optional<std::string> inserts_dir;
    // chrono::Local::now() is nondeterministic; tests inject a fixed clock.
    std::optional<std::string> fixed_hhmm, fixed_hhmmss;
    long max_iterations = 100000;       // ERR_LIMIT guard (the reference has none)
    size_t max_bytes = size_t(1) << 26;  // ERR_LIMIT guard
};

inline std::string replace_all(const std::string& s, const std::string& from, const std::string& to) {
    // Rust str::replace: non-overlapping matches, left to right.
    std::string out;
    out.reserve(s.size());
    size_t pos = 0;
    for (;;) {
        size_t f = s.find(from, pos);
        if (f == std::string::npos) { out.append(s, pos, std::string::npos); break; }
        out.append(s, pos, f - pos);
        out += to;
        pos = f + from.size();
    }
    return out;
}

inline size_t count_matches(const std::string& s, const std::string& pat) {
    // Rust str::matches(..).count(): non-overlapping
    size_t n = 0, pos = 0;
    while ((pos = s.find(pat, pos)) != std::string::npos) { ++n; pos += pat.size(); }
    return n;
}

static const std::string ESCAPED_START = std::string(1, ESCAPE) + INSERT_START;  // "\{"
static const std::string ESCAPED_STOP = std::string(1, ESCAPE) + INSERT_STOP;    // "\}"
static const std::string REPLACED_START = ".\xE3\x80\xA0";                         // ".〠" interp.rs:40
static const std::string REPLACED_STOP = "\xE3\x80\xA0.";                          // "〠." interp.rs:41

// interp.rs:314-322
inline std::string value_to_string(const Value& v) {
    if (v.raw) return v.s;
    switch (v.kind) {
        case Value::String: return v.s;
        case Value::Int: case Value::UInt: case Value::Float: return number_to_string(v);
        case Value::Bool: return v.b ? "true" : "false";
        case Value::Arr: { std::string out; for (auto& e : *v.a) out += value_to_string(e); return out; }
        default: return to_json(v);
    }
}

// interp.rs:11-29.  The reference walks `char`s; every char it tests is ASCII and UTF-8
// never embeds ASCII bytes in multi-byte sequences, so the byte walk is equivalent.
inline std::optional<std::string> get_simple_insertkey(const std::string& c) {
    long depth = 0;
    size_t n = c.size();
    if (n < 2 || c.front() != INSERT_START || c.back() != INSERT_STOP) return std::nullopt;
    for (size_t i = 0; i < n; ++i) {
        if (c[i] == INSERT_STOP) depth -= 1;
        if ((depth == 0) != (i == 0 || i == n - 1)) return std::nullopt;
        if (c[i] == INSERT_START) depth += 1;
    }
    return c.substr(1, n - 2);
}

inline Value recursive_escape(const Value& v);

inline bool path_exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline std::string trim_ws(const std::string& s) {
    // Rust str::trim(): Unicode White_Space; ASCII subset + U+0085/U+00A0 etc. are not handled
    // here beyond ASCII (inserts-dir files in the reference's examples are ASCII).
    size_t a = 0, b = s.size();
    auto ws = [](unsigned char ch) { return ch == ' ' || (ch >= 9 && ch <= 13); };
    while (a < b && ws((unsigned char)s[a])) ++a;
    while (b > a && ws((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

// interp.rs:91-137
inline Value get_interpdata(const Object& inserts, const std::string& key, const Ctx& ctx) {
    if (key == "HH:MM" || key == "HH:MM:SS") {
        bool secs = key.size() == 8;
        if (secs && ctx.fixed_hhmmss) return Value::string(*ctx.fixed_hhmmss);
        if (!secs && ctx.fixed_hhmm) return Value::string(*ctx.fixed_hhmm);
        std::time_t t = std::time(nullptr);
        std::tm tmv{};
        localtime_r(&t, &tmv);
        char buf[16];
        std::strftime(buf, sizeof buf, secs ? "%H:%M:%S" : "%H:%M", &tmv);
        return Value::string(buf);
    }
    if (key.empty()) throw InterpError{ERR_EMPTY, "", "Tried to interpolate empty string ''"};

    bool is_arg = key.rfind("ARG", 0) == 0;
    if (is_arg) for (size_t i = 3; i < key.size(); ++i) if (key[i] < '0' || key[i] > '9') { is_arg = false; break; }
    if (is_arg) {
        auto it = inserts.find(key);
        if (it != inserts.end()) return it->second.deep_clone();
        throw InterpError{ERR_ARG, key, "Argument interpolation key '" + key + "' is used but not provided"};
    }
    auto it = inserts.find(key);
    if (it != inserts.end()) return it->second.deep_clone();

    if (ctx.inserts_dir) {
        std::string j5 = *ctx.inserts_dir + "/" + key + ".json5";
        if (path_exists(j5)) {
            std::ifstream f(j5, std::ios::binary);
            if (!f) throw InterpError{ERR_IO, key, "cannot read " + j5};
            std::stringstream ss; ss << f.rdbuf();
            try { return recursive_escape(parse_json(ss.str())); }
            catch (const std::runtime_error& e) { throw InterpError{ERR_IO, key, e.what()}; }
        }
        std::string plain = *ctx.inserts_dir + "/" + key;
        if (path_exists(plain)) {
            std::ifstream f(plain, std::ios::binary);
            if (!f) throw InterpError{ERR_IO, key, "cannot read " + plain};
            std::stringstream ss; ss << f.rdbuf();
            return recursive_escape(Value::string(trim_ws(ss.str())));
        }
    }
    throw InterpError{ERR_NOT_FOUND, key, "Could not find variable '" + key + "'"};
}

// interp.rs:31-89
inline Value interpolate_inserts(const Object& inserts, const std::string& content, const Ctx& ctx, int depth = 0) {
    if (depth > 2000) throw InterpError{ERR_LIMIT, "", "expansion limit exceeded"};
    std::string s = content;
    s = replace_all(s, ESCAPED_START, REPLACED_START);  // :42
    s = replace_all(s, ESCAPED_STOP, REPLACED_STOP);    // :43

    if (auto insertkey = get_simple_insertkey(s)) {      // :45
        if (auto subkey = get_simple_insertkey(*insertkey)) {  // :46
            Value inner = interpolate_inserts(inserts, std::string(1, INSERT_START) + *subkey + INSERT_STOP, ctx, depth + 1);
            return get_interpdata(inserts, value_to_string(inner), ctx);
        }
        Value inner = interpolate_inserts(inserts, *insertkey, ctx, depth + 1);  // :50
        return get_interpdata(inserts, value_to_string(inner), ctx);              // :51
    }

    long iters = 0;
    while (s.find(INSERT_START) != std::string::npos) {  // :54
        if (++iters > ctx.max_iterations || s.size() > ctx.max_bytes)
            throw InterpError{ERR_LIMIT, "", "expansion limit exceeded"};
        size_t n_starts = count_matches(s, std::string(1, INSERT_START)) - count_matches(s, ESCAPED_START);  // :55
        size_t n_stops = count_matches(s, std::string(1, INSERT_STOP)) - count_matches(s, ESCAPED_STOP);     // :56
        if (n_starts != n_stops)
            throw InterpError{ERR_UNEVEN, s, "Interpolation error: uneven number of '{' and '}' in: " + s};  // :57-61
        size_t outer_from = s.rfind(INSERT_START);            // :62
        size_t inner_to = s.find(INSERT_STOP, outer_from + 1);  // :63-66
        if (inner_to == std::string::npos) throw InterpError{ERR_PANIC, "", "panic: called `Option::unwrap()` on a `None` value"};
        std::string inner = s.substr(outer_from + 1, inner_to - outer_from - 1);
        inner = replace_all(inner, REPLACED_START, ESCAPED_START);  // :68
        inner = replace_all(inner, REPLACED_STOP, ESCAPED_STOP);    // :69
        Value insert_value = get_interpdata(inserts, inner, ctx);   // :70
        std::string insert_str;
        if (insert_value.kind == Value::String) insert_str = insert_value.raw ? insert_value.s : insert_value.s;
        else if (insert_value.is_number()) insert_str = number_to_string(insert_value);
        else if (insert_value.kind == Value::Arr) insert_str = value_to_string(insert_value);  // :74 join("")
        else throw InterpError{ERR_UNSUPPORTED, inner, "Trying to interpolate '" + inner + "' of unsupported type"};
        s = s.substr(0, outer_from) + insert_str + s.substr(inner_to + 1);  // :81
        s = replace_all(s, ESCAPED_START, REPLACED_START);                  // :82
        s = replace_all(s, ESCAPED_STOP, REPLACED_STOP);                    // :83
    }
    s = replace_all(s, REPLACED_START, ESCAPED_START);  // :86
    s = replace_all(s, REPLACED_STOP, ESCAPED_STOP);    // :87
    return Value::string(std::move(s));
}

// interp.rs:147-161
inline std::string unescape_str(const std::string& s) {
    return replace_all(replace_all(s, ESCAPED_START, std::string(1, INSERT_START)), ESCAPED_STOP, std::string(1, INSERT_STOP));
}
inline Value recursive_unescape(const Value& v) {
    switch (v.kind) {
        case Value::String: return Value::string(unescape_str(v.s));
        case Value::Arr: { Array a; for (auto& e : *v.a) a.push_back(recursive_unescape(e)); return Value::array(std::move(a)); }
        case Value::Obj: { Object o; for (auto& kv : *v.o) o[unescape_str(kv.first)] = recursive_unescape(kv.second); return Value::object(std::move(o)); }
        default: return v;
    }
}
// interp.rs:163-177
inline std::string escape_str(const std::string& s) {
    return replace_all(replace_all(s, std::string(1, INSERT_START), ESCAPED_START), std::string(1, INSERT_STOP), ESCAPED_STOP);
}
inline Value recursive_escape(const Value& v) {
    switch (v.kind) {
        case Value::String: return Value::string(escape_str(v.s));
        case Value::Arr: { Array a; for (auto& e : *v.a) a.push_back(recursive_escape(e)); return Value::array(std::move(a)); }
        case Value::Obj: { Object o; for (auto& kv : *v.o) o[escape_str(kv.first)] = recursive_escape(kv.second); return Value::object(std::move(o)); }
        default: return v;
    }
}

// interp.rs:179-246
inline Value recursive_interpolate(const Object& inserts, const Value& value, const Ctx& ctx) {
    if (value.kind == Value::String) {
        if (auto insertkey = get_simple_insertkey(value.s)) {  // :184-196
            try { return interpolate_inserts(inserts, std::string(1, INSERT_START) + *insertkey + INSERT_STOP, ctx); }
            catch (const InterpError& e) { if (e.code == ERR_PANIC || e.code == ERR_LIMIT) throw; return Value::string(value.s); }
        }
        try { return interpolate_inserts(inserts, value.s, ctx); }  // :199-202
        catch (const InterpError& e) { if (e.code == ERR_PANIC || e.code == ERR_LIMIT) throw; return Value::string(value.s); }
    }
    if (value.kind == Value::Arr) {
        Array a;
        for (auto& e : *value.a) a.push_back(recursive_interpolate(inserts, e, ctx));
        return Value::array(std::move(a));
    }
    if (value.kind == Value::Obj) {
        const Object& obj = *value.o;
        auto cit = obj.find("cmd");
        if (cit != obj.end() && cit->second.kind == Value::String) {
            const std::string& cmd = cit->second.s;
            if (cmd == "goto_map" || cmd == "replace_map") return value.deep_clone();  // :210-212
            if (cmd == "for" || cmd == "serial" || cmd == "parallel_wait" || cmd == "parallel_race") {  // :213-233
                Value out = value.deep_clone();
                auto tit = out.o->find("tasks");
                if (tit != out.o->end()) {
                    Value& tv = tit->second;
                    if (tv.kind == Value::String) {
                        if (auto k = get_simple_insertkey(tv.s)) tv = get_interpdata(inserts, *k, ctx);
                    } else if (tv.kind == Value::Arr) {
                        for (auto& e : *tv.a)
                            if (e.kind == Value::String)
                                if (auto k = get_simple_insertkey(e.s)) e = get_interpdata(inserts, *k, ctx);
                    }
                }
                return out;
            }
        }
        Object out;
        for (auto& kv : obj) {  // sorted order; later duplicates win (:235-242)
            Value nk = recursive_interpolate(inserts, Value::string(kv.first), ctx);
            Value nv = recursive_interpolate(inserts, kv.second, ctx);
            out[value_to_string(nk)] = std::move(nv);
        }
        return Value::object(std::move(out));
    }
    return value;
}

// The resolver calls recursive_interpolate issues for `value`, in its traversal order (interp.rs:179-246): every string
// and every object key goes to interpolate_inserts (:186 / :199), the simple keys of a control task's `tasks` to
// get_interpdata (:217, :224); goto_map / replace_map objects and the bodies of control tasks are not entered.
// This is the batch a replayed program presents to the resolver (BASELINE.json configs 1-3).
inline void interpolation_trace(const Value& value, std::vector<std::string>& templates, std::vector<std::string>& lookups) {
    if (value.kind == Value::String) { templates.push_back(value.s); return; }
    if (value.kind == Value::Arr) { for (auto& e : *value.a) interpolation_trace(e, templates, lookups); return; }
    if (value.kind != Value::Obj) return;
    const Object& obj = *value.o;
    auto cit = obj.find("cmd");
    if (cit != obj.end() && cit->second.kind == Value::String) {
        const std::string& cmd = cit->second.s;
        if (cmd == "goto_map" || cmd == "replace_map") return;
        if (cmd == "for" || cmd == "serial" || cmd == "parallel_wait" || cmd == "parallel_race") {
            auto tit = obj.find("tasks");
            if (tit != obj.end()) {
                const Value& tv = tit->second;
                if (tv.kind == Value::String) { if (auto k = get_simple_insertkey(tv.s)) lookups.push_back(*k); }
                else if (tv.kind == Value::Arr)
                    for (auto& e : *tv.a)
                        if (e.kind == Value::String)
                            if (auto k = get_simple_insertkey(e.s)) lookups.push_back(*k);
            }
            return;
        }
    }
    for (auto& kv : obj) { templates.push_back(kv.first); interpolation_trace(kv.second, templates, lookups); }
}

// interp.rs:273-312
inline void extract_from_str(const std::string& s, std::vector<std::string>& keys) {
    long depth = 0;
    std::string current;
    bool in_key = false, escaped = false;
    // The reference iterates chars and pushes whole chars; pushing the bytes of a multi-byte
    // char one at a time yields the same string because none of them equals an ASCII delimiter.
    for (char ch : s) {
        if (escaped) { escaped = false; if (in_key) current.push_back(ch); continue; }
        if (ch == ESCAPE) { escaped = true; continue; }
        if (ch == INSERT_START) {
            depth += 1;
            if (depth == 1) { in_key = true; current.clear(); continue; }
        }
        if (ch == INSERT_STOP) {
            if (depth == 1 && in_key) { keys.push_back(current); in_key = false; depth -= 1; continue; }
            if (depth > 0) depth -= 1;
        }
        if (in_key) current.push_back(ch);
    }
}
// interp.rs:248-271
inline void extract_insert_keys(const Value& v, std::vector<std::string>& keys) {
    switch (v.kind) {
        case Value::String: extract_from_str(v.s, keys); break;
        case Value::Arr: for (auto& e : *v.a) extract_insert_keys(e, keys); break;
        case Value::Obj: for (auto& kv : *v.o) { extract_from_str(kv.first, keys); extract_insert_keys(kv.second, keys); } break;
        default: break;
    }
}

// runtime.rs:1633-1647.  The reference compiles "^" + (".*" | escaped literal)... + "$" with
// dot_matches_new_line per call.  `regex` 1.10 is not vendored; the anchored regex it builds
// recognises exactly the language below (literal runs separated by arbitrary gaps), so a
// direct matcher restates it.  Matching bytes instead of chars is equivalent on valid UTF-8.
// The crate's compiled-size limit (=> build error => false) is PARITY UNPINNED.
inline bool wildcard_match(const std::string& p, const std::string& s) {
    size_t pi = 0, si = 0, star = std::string::npos, mark = 0;
    while (si < s.size()) {
        if (pi < p.size() && p[pi] != '*' && p[pi] == s[si]) { ++pi; ++si; }
        else if (pi < p.size() && p[pi] == '*') { star = pi++; mark = si; }
        else if (star != std::string::npos) { pi = star + 1; si = ++mark; }
        else return false;
    }
    while (pi < p.size() && p[pi] == '*') ++pi;
    return pi == p.size();
}

// runtime.rs:1754-1775: "^" + ("(.*)" | escaped literal)... + "$"; captures of the leftmost
// match under greedy backtracking semantics (each `*` takes the longest run that still lets
// the rest match, decided left to right).
inline std::vector<std::string> wildcard_captures(const std::string& p, const std::string& s) {
    std::vector<std::string> lits(1);
    for (char c : p) { if (c == '*') lits.emplace_back(); else lits.back().push_back(c); }
    size_t m = lits.size() - 1;  // number of stars
    std::vector<std::string> caps;
    if (!wildcard_match(p, s)) return caps;
    if (m == 0) return caps;  // filter_map over zero groups
    // memoised feasibility: can lits[j..] with stars between them match s[pos..] when
    // literal j starts exactly at pos?
    size_t n = s.size();
    std::vector<std::vector<int8_t>> memo(lits.size(), std::vector<int8_t>(n + 1, -1));
    std::function<bool(size_t, size_t)> ok = [&](size_t j, size_t pos) -> bool {
        if (pos > n) return false;
        int8_t& mm = memo[j][pos];
        if (mm >= 0) return mm;
        bool r = false;
        const std::string& L = lits[j];
        if (pos + L.size() <= n && s.compare(pos, L.size(), L) == 0) {
            size_t after = pos + L.size();
            if (j == m) r = (after == n);
            else for (size_t e = n + 1; e-- > after;) if (ok(j + 1, e)) { r = true; break; }
        }
        mm = r;
        return r;
    };
    size_t pos = lits[0].size();
    for (size_t j = 1; j <= m; ++j) {
        // star j-1 spans [pos, e): greedy = largest e for which the rest still matches
        size_t e = n + 1;
        while (e-- > pos) if (ok(j, e)) break;
        caps.push_back(s.substr(pos, e - pos));
        pos = e + lits[j].size();
    }
    return caps;
}

// runtime.rs:1198-1218 (delete) and 1219-1239 (delete_except): keys snapshot in sorted order,
// remove when (any wildcard matches) != invert; returns the deleted keys in that order.
inline std::vector<std::string> delete_matching(Object& inserts, const std::vector<Value>& wildcards, bool except) {
    std::vector<std::string> keys, deleted;
    for (auto& kv : inserts) keys.push_back(kv.first);
    for (auto& k : keys) {
        bool any = false;
        for (auto& w : wildcards) if (wildcard_match(value_to_string(w), k)) { any = true; break; }
        if (any != except) { inserts.erase(k); deleted.push_back(k); }
    }
    return deleted;
}

// ---- "next" rows of SURVEY.md §8(f): the two callers that run the resolver in a loop --------------------------

// anyhow errors of runtime.rs that are not resolver errors (code -2: the message is the whole contract)
inline InterpError task_error(const std::string& msg) { return InterpError{-2, "", msg}; }

// runtime.rs:1733-1752
inline std::optional<Value> find_null_map_value(const Array& maps, const Object& inserts, const Ctx& ctx) {
    for (auto& map : maps) {
        if (map.kind != Value::Obj) continue;
        for (auto& kv : *map.o) {
            if (kv.first == "NULL") return kv.second.deep_clone();
            if (kv.first.find('{') != std::string::npos) {
                try {
                    if (value_to_string(interpolate_inserts(inserts, kv.first, ctx)) == "NULL") return kv.second.deep_clone();
                } catch (const InterpError& e) { if (e.code == ERR_PANIC) throw; }  // `if let Ok(..)` does not catch the unwrap panic of interp.rs:63-66
            }
        }
    }
    return std::nullopt;
}

// runtime.rs:1658-1692 (the nested fn replace_str)
inline std::string replace_str(std::string text, const Array& maps, const Object& inserts, const Ctx& ctx, bool repeat_until_done) {
    std::set<std::string> seen;  // a text that comes back can only repeat its cycle
    for (long guard = 0;; ++guard) {
        if (guard > 10000 || !seen.insert(text).second) throw InterpError{ERR_LIMIT, "", "expansion limit exceeded"};  // the reference would spin forever
        const std::string current = value_to_string(interpolate_inserts(inserts, text, ctx));
        std::optional<std::string> replaced;
        for (auto& map : maps) {
            if (map.kind != Value::Obj) throw task_error("replace_map expects object");
            if (map.o->empty()) throw task_error("replace_map entry empty");
            const auto& kv = *map.o->begin();  // obj.iter().next(): first key in sorted order
            const std::string key = value_to_string(interpolate_inserts(inserts, kv.first, ctx));
            if (wildcard_match(key, current)) {
                const std::vector<std::string> caps = wildcard_captures(key, current);
                Object extra = inserts;
                for (size_t i = 0; i < caps.size(); ++i) extra[std::to_string(i + 1)] = Value::string(caps[i]);
                replaced = value_to_string(interpolate_inserts(extra, kv.second.is_string() ? kv.second.s : std::string(), ctx));
                break;
            }
        }
        const std::string new_text = replaced ? *replaced : current;
        if (!repeat_until_done || new_text == text) return new_text;
        text = new_text;
    }
}

// runtime.rs:1649-1731.  Every `?` inside the match arms returns from the function, so the trailing
// "NULL handler" match only ever sees Ok: errors propagate, except for the simple-key shortcut at :1696-1702.
inline Value replace_map(const Value& item, const Array& maps, const Object& inserts, const Ctx& ctx, bool repeat_until_done) {
    const std::optional<Value> null_value = find_null_map_value(maps, inserts, ctx);
    if (item.kind == Value::String) {
        if (get_simple_insertkey(item.s) && null_value) {
            bool failed = false;
            try { interpolate_inserts(inserts, item.s, ctx); } catch (const InterpError& e) { if (e.code == ERR_PANIC) throw; failed = true; }
            if (failed) return *null_value;
        }
        return Value::string(replace_str(item.s, maps, inserts, ctx, repeat_until_done));
    }
    if (item.kind == Value::Arr) {
        Array out;
        for (auto& v : *item.a) out.push_back(replace_map(v, maps, inserts, ctx, repeat_until_done));
        return Value::array(std::move(out));
    }
    if (item.kind == Value::Obj) {
        Object out;
        for (auto& kv : *item.o) {
            std::string nk = replace_str(kv.first, maps, inserts, ctx, repeat_until_done);
            out[nk] = replace_map(kv.second, maps, inserts, ctx, repeat_until_done);  // Map::insert: a later duplicate wins
        }
        return Value::object(std::move(out));
    }
    return item.deep_clone();
}

// runtime.rs:1085-1133: the target a goto_map task jumps to ("CONTINUE" = fall through); plus what it logs.
struct GotoChoice { std::string value_text, target; bool interp_error; };
inline GotoChoice goto_map(const std::string& text, const Array& target_maps, const Object& inserts, const Ctx& ctx) {
    GotoChoice g{"", "", false};
    try { g.value_text = value_to_string(interpolate_inserts(inserts, text, ctx)); }
    catch (const InterpError& e) { if (e.code == ERR_PANIC) throw; g.interp_error = true; g.value_text = "NULL"; }
    std::optional<std::string> target;
    for (auto& entry : target_maps) {
        if (entry.kind != Value::Obj) throw task_error("target_maps entry must be object");
        if (entry.o->empty()) throw task_error("target_maps entry empty");
        const auto& kv = *entry.o->begin();
        const std::string key = value_to_string(interpolate_inserts(inserts, kv.first, ctx));
        const std::string vtxt = kv.second.is_string() ? kv.second.s : std::string();
        if (g.interp_error) {
            if (key == "NULL") { target = value_to_string(interpolate_inserts(inserts, vtxt, ctx)); break; }
        } else {
            const std::string val = value_to_string(interpolate_inserts(inserts, vtxt, ctx));
            if (wildcard_match(key, g.value_text)) { target = val; break; }
        }
    }
    if (!target) {
        if (g.interp_error) throw task_error("goto_map value could not be resolved but 'NULL' is not a key in target_maps");
        throw task_error("goto_map has no matches for '" + g.value_text + "'");
    }
    g.target = *target;
    return g;
}

// ---- program loader: rust-project/src/parser.rs:8-93 -----------------------------------------
// add_line_numbers (parser.rs:74-93): per line, every non-overlapping leftmost match of
//   (\bcmd\b|"cmd"|'cmd')\s*:\s*("([^"\\]|\\.)*"|'([^'\\]|\\.)*')(\s*(?:,|\}))
// becomes  key:val, line:N trail .  \b and \s are taken over ASCII (bytes >= 0x80 count as word characters);
// the regex crate's Unicode classes are PARITY UNPINNED.
inline bool lp_word(unsigned char c) { return std::isalnum(c) || c == '_' || c >= 0x80; }
inline bool lp_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
inline std::string add_line_numbers(const std::string& input) {
    std::string out;
    size_t line_no = 0, at = 0;
    while (at < input.size()) {  // str::lines(): split on \n, a trailing \r is dropped, no empty last line
        size_t nl = input.find('\n', at);
        std::string line = input.substr(at, nl == std::string::npos ? std::string::npos : nl - at);
        at = nl == std::string::npos ? input.size() : nl + 1;
        if (nl != std::string::npos && !line.empty() && line.back() == '\r') line.pop_back();
        ++line_no;
        size_t i = 0, copied = 0;
        std::string res;
        const size_t n = line.size();
        while (i < n) {
            size_t k = std::string::npos;  // end of the key alternative matched at i
            if (line.compare(i, 3, "cmd") == 0 && (i == 0 || !lp_word((unsigned char)line[i - 1])) &&
                (i + 3 == n || !lp_word((unsigned char)line[i + 3]))) k = i + 3;
            else if (line.compare(i, 5, "\"cmd\"") == 0 || line.compare(i, 5, "'cmd'") == 0) k = i + 5;
            if (k != std::string::npos) {
                size_t j = k;
                while (j < n && lp_space((unsigned char)line[j])) ++j;
                if (j < n && line[j] == ':') {
                    ++j;
                    while (j < n && lp_space((unsigned char)line[j])) ++j;
                    if (j < n && (line[j] == '"' || line[j] == '\'')) {
                        const char q = line[j];
                        size_t v = j + 1;
                        bool closed = false;
                        while (v < n) {
                            if (line[v] == '\\') { if (v + 1 >= n) break; v += 2; continue; }
                            if (line[v] == q) { closed = true; break; }
                            ++v;
                        }
                        if (closed) {
                            size_t t = v + 1;
                            while (t < n && lp_space((unsigned char)line[t])) ++t;
                            if (t < n && (line[t] == ',' || line[t] == '}')) {
                                res.append(line, copied, i - copied);
                                res.append(line, i, k - i).append(":").append(line, j, v + 1 - j);
                                res.append(", line:").append(std::to_string(line_no)).append(line, v + 1, t + 1 - (v + 1));
                                i = copied = t + 1;
                                continue;
                            }
                        }
                    }
                }
            }
            ++i;
        }
        res.append(line, copied, std::string::npos);
        out += res;
        out += '\n';
    }
    return out;
}

// parser.rs:8-64: the Program as an object {default_state, order, named_tasks, save_states, completion_args}
inline Value load_program(const std::string& raw) {
    Value root = parse_json(add_line_numbers(raw));  // json5::from_str: the JSON5 subset of oracle_value.hpp
    if (root.kind != Value::Obj) throw task_error("Program root must be an object");
    Object& obj = *root.o;
    if (!obj.count("named_tasks") && obj.count("tasks")) { obj["named_tasks"] = obj["tasks"]; obj.erase("tasks"); }
    auto need_obj = [&](const char* k) -> const Object& {
        auto it = obj.find(k);
        if (it == obj.end() || it->second.kind != Value::Obj) throw task_error(std::string("Program missing '") + k + "' object");
        return *it->second.o;
    };
    auto as_task = [](const Value& v) {
        if (v.kind == Value::Obj) return;
        std::string j;
        to_json(v, j);  // the reference prints the Value's Debug form ({value:?}): compact JSON here, PARITY UNPINNED
        throw task_error("Task must be an object, got " + j);
    };
    Object out;
    out["default_state"] = Value::object(need_obj("default_state"));
    auto ord = obj.find("order");
    if (ord == obj.end() || ord->second.kind != Value::Arr) throw task_error("Program missing 'order' array");
    for (auto& t : *ord->second.a) as_task(t);
    out["order"] = ord->second;
    const Object& named = need_obj("named_tasks");
    for (auto& kv : named) as_task(kv.second);
    out["named_tasks"] = Value::object(named);
    out["save_states"] = Value::object(need_obj("save_states"));
    auto ca = obj.find("completion_args");
    out["completion_args"] = (ca != obj.end() && ca->second.kind == Value::Obj) ? ca->second : Value::object();
    return Value::object(std::move(out));
}

}  // namespace orc
