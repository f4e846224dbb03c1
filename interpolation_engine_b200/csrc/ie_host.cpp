// Host mirror of the reference's resolver API over the CUDA batch ABI.
//
// Function names, argument meaning and error texts follow rust-project/src/interp.rs; the tree
// walkers gather every string they would resolve into ONE batch per call (the shape of
// runtime.rs:700-701, which resolves a whole task object against a snapshot of the inserts), run it
// through ie_resolve_batch / ie_escape_batch / ie_glob_sweep on the GPU and rebuild the tree.
// Nothing here interprets "{...}" text on the CPU; the host only does what the reference's leaves
// do outside the resolver proper: JSON (de)serialisation, value_to_string, the wall clock
// (interp.rs:96-104), the --inserts-dir file reads (interp.rs:122-134) and message formatting.
#include "ie_host.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <map>
#include <mutex>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>

namespace ie_host {

namespace {

// ---- JSON value ------------------------------------------------------------------------------
struct JVal;
using JArr = std::vector<JVal>;
using JObj = std::map<std::string, JVal>;  // serde_json::Map without preserve_order = BTreeMap

struct JVal {
    enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false;
    std::string s;  // Str: text; Num: canonical rendering (serde_json::Number::to_string)
    std::shared_ptr<JArr> a;
    std::shared_ptr<JObj> o;
    static JVal str(std::string x) { JVal v; v.t = Str; v.s = std::move(x); return v; }
    static JVal num(std::string x) { JVal v; v.t = Num; v.s = std::move(x); return v; }
    static JVal boolean(bool x) { JVal v; v.t = Bool; v.b = x; return v; }
    static JVal arr(JArr x = {}) { JVal v; v.t = Arr; v.a = std::make_shared<JArr>(std::move(x)); return v; }
    static JVal obj(JObj x = {}) { JVal v; v.t = Obj; v.o = std::make_shared<JObj>(std::move(x)); return v; }
};

// serde_json renders f64 through ryu's "pretty" layout (not vendored: PARITY UNPINNED)
std::string render_f64(double x) {
    if (!std::isfinite(x)) return "null";
    if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(x), std::chars_format::scientific);
    std::string sci(buf, r.ptr);
    const size_t epos = sci.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; ++i) if (sci[i] != '.') digits.push_back(sci[i]);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int e10 = std::atoi(sci.c_str() + epos + 1);
    const int len = (int)digits.size();
    const int kk = e10 + 1;  // decimal point position relative to the digit string
    const int k = kk - len;
    std::string out = x < 0 ? "-" : "";
    if (0 <= k && kk <= 16) out += digits + std::string((size_t)k, '0') + ".0";
    else if (0 < kk && kk <= 16) out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    else if (-5 < kk && kk <= 0) out += "0." + std::string((size_t)(-kk), '0') + digits;
    else if (len == 1) out += digits + "e" + std::to_string(kk - 1);
    else out += digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
    return out;
}

struct Reader {
    const char* p; const char* e;
    [[noreturn]] void bad(const char* m) { throw std::runtime_error(std::string("JSON parse error: ") + m); }
    void skip() {
        for (;;) {
            while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
            if (e - p >= 2 && p[0] == '/' && p[1] == '/') { while (p < e && *p != '\n') ++p; }
            else if (e - p >= 2 && p[0] == '/' && p[1] == '*') { p += 2; while (e - p >= 2 && !(p[0] == '*' && p[1] == '/')) ++p; p = (e - p >= 2) ? p + 2 : e; }
            else return;
        }
    }
    static void utf8(uint32_t c, std::string& o) {
        if (c < 0x80) o += (char)c;
        else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 63)); }
        else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
        else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 63)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
    }
    uint32_t hex4() {
        if (e - p < 4) bad("\\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i, ++p) {
            const char c = *p;
            v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : (bad("hex digit"), 0));
        }
        return v;
    }
    std::string string_lit() {
        const char q = *p++;
        std::string o;
        for (;;) {
            if (p >= e) bad("unterminated string");
            const char c = *p++;
            if (c == q) return o;
            if (c != '\\') { o += c; continue; }
            if (p >= e) bad("escape");
            const char n = *p++;
            switch (n) {
                case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
                case 'b': o += '\b'; break; case 'f': o += '\f'; break; case '0': o += '\0'; break;
                case '\n': break;
                case 'u': {
                    uint32_t c1 = hex4();
                    if (c1 >= 0xD800 && c1 < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') { p += 2; c1 = 0x10000 + ((c1 - 0xD800) << 10) + (hex4() - 0xDC00); }
                    utf8(c1, o);
                    break;
                }
                default: o += n;
            }
        }
    }
    JVal number_lit() {
        const char* b = p;
        if (p < e && (*p == '-' || *p == '+')) ++p;
        bool flt = false;
        while (p < e && (std::isdigit((unsigned char)*p) || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
            flt |= (*p == '.' || *p == 'e' || *p == 'E');
            ++p;
        }
        std::string tok(b, p);
        if (!tok.empty() && tok[0] == '+') tok.erase(0, 1);
        if (tok.empty() || tok == "-") bad("number");
        if (!flt) {
            // canonical integer text: strip leading zeros; fall back to f64 beyond u64/i64 like serde_json
            const bool neg = tok[0] == '-';
            size_t i = neg ? 1 : 0;
            while (i + 1 < tok.size() && tok[i] == '0') ++i;
            const std::string mag = tok.substr(i);
            const std::string lim = neg ? "9223372036854775808" : "18446744073709551615";
            if (mag.size() < lim.size() || (mag.size() == lim.size() && mag <= lim)) return JVal::num((neg && mag != "0" ? "-" : "") + mag);
        }
        return JVal::num(render_f64(std::strtod(tok.c_str(), nullptr)));
    }
    JVal value() {
        skip();
        if (p >= e) bad("unexpected end");
        const char c = *p;
        if (c == '{') {
            ++p;
            JObj o;
            for (;;) {
                skip();
                if (p < e && *p == '}') { ++p; break; }
                std::string k;
                if (p < e && (*p == '"' || *p == '\'')) k = string_lit();
                else { const char* b = p; while (p < e && (std::isalnum((unsigned char)*p) || *p == '_' || *p == '$')) ++p; if (p == b) bad("object key"); k.assign(b, p); }
                skip();
                if (p >= e || *p != ':') bad("':' expected");
                ++p;
                o[k] = value();
                skip();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; break; }
                bad("',' or '}' expected");
            }
            return JVal::obj(std::move(o));
        }
        if (c == '[') {
            ++p;
            JArr a;
            for (;;) {
                skip();
                if (p < e && *p == ']') { ++p; break; }
                a.push_back(value());
                skip();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; break; }
                bad("',' or ']' expected");
            }
            return JVal::arr(std::move(a));
        }
        if (c == '"' || c == '\'') return JVal::str(string_lit());
        if (e - p >= 4 && !std::memcmp(p, "true", 4)) { p += 4; return JVal::boolean(true); }
        if (e - p >= 5 && !std::memcmp(p, "false", 5)) { p += 5; return JVal::boolean(false); }
        if (e - p >= 4 && !std::memcmp(p, "null", 4)) { p += 4; return JVal(); }
        return number_lit();
    }
};

JVal parse(const std::string& text) {
    Reader r{text.data(), text.data() + text.size()};
    JVal v = r.value();
    r.skip();
    if (r.p != r.e) r.bad("trailing characters");
    return v;
}

void quote(const std::string& s, std::string& o) {
    static const char* hexd = "0123456789abcdef";
    o += '"';
    for (unsigned char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '\n') o += "\\n";
        else if (c == '\r') o += "\\r";
        else if (c == '\t') o += "\\t";
        else if (c == '\b') o += "\\b";
        else if (c == '\f') o += "\\f";
        else if (c < 0x20) { o += "\\u00"; o += hexd[c >> 4]; o += hexd[c & 15]; }
        else o += (char)c;
    }
    o += '"';
}
void dump(const JVal& v, std::string& o) {  // serde_json::to_string (compact)
    switch (v.t) {
        case JVal::Null: o += "null"; break;
        case JVal::Bool: o += v.b ? "true" : "false"; break;
        case JVal::Num: o += v.s; break;
        case JVal::Str: quote(v.s, o); break;
        case JVal::Arr: { o += '['; bool f = true; for (auto& x : *v.a) { if (!f) o += ','; f = false; dump(x, o); } o += ']'; break; }
        case JVal::Obj: { o += '{'; bool f = true; for (auto& kv : *v.o) { if (!f) o += ','; f = false; quote(kv.first, o); o += ':'; dump(kv.second, o); } o += '}'; break; }
    }
}

// interp.rs:314-322
void render(const JVal& v, std::string& o) {
    switch (v.t) {
        case JVal::Str: case JVal::Num: o += v.s; break;
        case JVal::Bool: o += v.b ? "true" : "false"; break;
        case JVal::Arr: for (auto& x : *v.a) render(x, o); break;
        default: dump(v, o);
    }
}
std::string value_to_string(const JVal& v) { std::string o; render(v, o); return o; }
uint8_t tag_of(const JVal& v) {
    switch (v.t) {
        case JVal::Null: return IE_TAG_NULL; case JVal::Bool: return IE_TAG_BOOL; case JVal::Num: return IE_TAG_NUMBER;
        case JVal::Str: return IE_TAG_STRING; case JVal::Arr: return IE_TAG_ARRAY; default: return IE_TAG_OBJECT;
    }
}

struct Arena {
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> offs{0};
    void push(const std::string& s) { bytes.insert(bytes.end(), s.begin(), s.end()); offs.push_back(bytes.size()); }
    uint64_t n() const { return offs.size() - 1; }
    const uint8_t* data() const { static const uint8_t z = 0; return bytes.empty() ? &z : bytes.data(); }
};

struct ApiError {
    int code;
    std::string message, payload;
};
struct CallFailure {  // engine-level failure (CUDA etc.): aborts the call
    ie_status_t st;
    std::string why;
};
void check(ie_status_t st) { if (st != IE_OK) throw CallFailure{st, ie_last_error()}; }

std::string message_for(int code, const std::string& payload) {
    switch (code) {
        case IE_RES_UNEVEN: return "Interpolation error: uneven number of '{' and '}' in: " + payload;  // interp.rs:58-60
        case IE_RES_UNSUPPORTED: return "Trying to interpolate '" + payload + "' of unsupported type";    // :76-78
        case IE_RES_EMPTY_KEY: return "Tried to interpolate empty string ''";                             // :105
        case IE_RES_ARG_MISSING: return "Argument interpolation key '" + payload + "' is used but not provided";  // :113-115
        case IE_RES_NOT_FOUND: return "Could not find variable '" + payload + "'";                        // :136
        case IE_RES_PANIC: return "panic: called `Option::unwrap()` on a `None` value";                   // :66
        case IE_RES_LIMIT: return "expansion limit exceeded";
        default: return "error";
    }
}

// interp.rs:11-29 (pure text predicate used by the analyzer and the tree walker; not a resolve step)
bool simple_insertkey(const std::string& c, std::string* inner) {
    const size_t n = c.size();
    if (n < 2 || c.front() != '{' || c.back() != '}') return false;
    long depth = 0;
    for (size_t i = 0; i < n; ++i) {
        if (c[i] == '}') --depth;
        if ((depth == 0) != (i == 0 || i == n - 1)) return false;
        if (c[i] == '{') ++depth;
    }
    if (inner) *inner = c.substr(1, n - 2);
    return true;
}

bool is_arg_key(const std::string& k) {
    if (k.rfind("ARG", 0) != 0) return false;
    for (size_t i = 3; i < k.size(); ++i) if (k[i] < '0' || k[i] > '9') return false;
    return true;
}

// ---- a resolver session: one inserts snapshot packed on the device ------------------------------
struct Outcome {
    int code = 0;          // IE_RES_*
    std::string bytes;     // result text / error payload
    JVal typed;            // IE_RES_TYPED: the insert's value
    std::string io_error;  // non-empty: inserts-dir read/parse failure for this template's key
};

struct Session {
    ie_engine* e;
    JObj inserts;  // snapshot (plus inserts-dir hits, which the reference reads on a map miss)
    std::string hhmm, hhmmss, inserts_dir;
    bool has_dir = false;
    ie_table* table = nullptr;
    // entry ids the device reports for typed results: [0, n0) = the inserts as packed, n0 / n0 + 1 = the clock keys,
    // n0 + 2 + k = the k-th key set after the pack (set_interpdata on the packed table, ie_table_set)
    std::vector<const JVal*> entry_vals, extra_vals;
    std::map<std::string, uint32_t> entry_ids;
    std::string packed_hhmm, packed_hhmmss;
    std::map<std::string, std::string> dir_errors;
    std::map<std::string, bool> dir_probed;

    Session(ie_engine* eng, const JVal& args) : e(eng) {
        if (args.t == JVal::Obj) {
            auto it = args.o->find("inserts");
            if (it != args.o->end() && it->second.t == JVal::Obj) inserts = *it->second.o;
            auto d = args.o->find("inserts_dir");
            if (d != args.o->end() && d->second.t == JVal::Str) { inserts_dir = d->second.s; has_dir = true; }
        }
        read_clock(args, &hhmm, &hhmmss);
    }
    // the "clock" argument of a call (tests inject a fixed one), else chrono::Local::now() (interp.rs:98,102), once per call
    static void read_clock(const JVal& args, std::string* hm, std::string* hs) {
        hm->clear();
        hs->clear();
        if (args.t == JVal::Obj) {
            auto c = args.o->find("clock");
            if (c != args.o->end() && c->second.t == JVal::Obj) {
                auto a = c->second.o->find("hhmm"); if (a != c->second.o->end()) *hm = a->second.s;
                auto b = c->second.o->find("hhmmss"); if (b != c->second.o->end()) *hs = b->second.s;
            }
        }
        if (hm->empty() || hs->empty()) {
            std::time_t now = std::time(nullptr);
            std::tm tmv{};
            localtime_r(&now, &tmv);
            char buf[16];
            if (hm->empty()) { std::strftime(buf, sizeof buf, "%H:%M", &tmv); *hm = buf; }
            if (hs->empty()) { std::strftime(buf, sizeof buf, "%H:%M:%S", &tmv); *hs = buf; }
        }
    }
    ~Session() { if (table) ie_table_free(table); }

    void pack() {
        if (table) { ie_table_free(table); table = nullptr; }
        Arena keys, vals;
        std::vector<uint8_t> tags;
        entry_vals.clear();
        extra_vals.clear();
        entry_ids.clear();
        for (auto& kv : inserts) {
            keys.push(kv.first);
            vals.push(value_to_string(kv.second));
            tags.push_back(tag_of(kv.second));
            entry_ids[kv.first] = (uint32_t)entry_vals.size();
            entry_vals.push_back(&kv.second);
        }
        static const uint8_t z = 0;
        check(ie_table_pack(e, keys.n(), keys.data(), keys.offs.data(), vals.data(), vals.offs.data(), tags.empty() ? &z : tags.data(),
                            hhmm.c_str(), hhmmss.c_str(), &table));
        packed_hhmm = hhmm;
        packed_hhmmss = hhmmss;
    }

    JVal entry_value(uint32_t entry) const {
        const uint32_t n0 = (uint32_t)entry_vals.size();
        if (entry < n0) return *entry_vals[entry];
        if (entry < n0 + 2) return JVal::str(entry == n0 ? hhmm : hhmmss);
        return *extra_vals[entry - n0 - 2];
    }

    // set_interpdata (interp.rs:139-141): the map and, when a table is packed, the table in place (ie_table_set); a table
    // without room for the operation is packed again.
    void set(const std::string& key, const JVal& value) {
        JVal& slot = inserts[key];
        slot = value;  // (map nodes do not move: entry_vals / extra_vals keep pointing at live values)
        if (!table || key == "HH:MM" || key == "HH:MM:SS") return;  // (the table's clock entries shadow inserts of those names)
        auto it = entry_ids.find(key);
        uint32_t id;
        if (it != entry_ids.end()) id = it->second;
        else {
            id = (uint32_t)(entry_vals.size() + 2 + extra_vals.size());
            extra_vals.push_back(&slot);
            entry_ids[key] = id;
        }
        const std::string text = value_to_string(value);
        const uint64_t ko[2] = {0, key.size()}, vo[2] = {0, text.size()};
        const uint8_t tag = tag_of(value);
        const ie_status_t st = ie_table_set(e, table, 0, 1, (const uint8_t*)key.data(), ko, (const uint8_t*)text.data(), vo, &tag, &id);
        if (st == IE_E_OVERFLOW) pack();
        else check(st);
    }
    // delete_interpdata (interp.rs:143-145)
    void del(const std::string& key) {
        if (!inserts.erase(key)) return;
        entry_ids.erase(key);
        if (!table || key == "HH:MM" || key == "HH:MM:SS") return;
        const uint64_t ko[2] = {0, key.size()};
        check(ie_table_delete(e, table, 0, 1, (const uint8_t*)key.data(), ko));
    }
    // A session that outlives a call: the clock keys of its table follow the wall clock (interp.rs:96-104 reads it per
    // lookup; here it is read once per call and patched in place when its rendering changed).
    void refresh_clock(const std::string& new_hhmm, const std::string& new_hhmmss) {
        hhmm = new_hhmm;
        hhmmss = new_hhmmss;
        if (!table || (hhmm == packed_hhmm && hhmmss == packed_hhmmss)) return;
        const std::string keys = "HH:MMHH:MM:SS", vals = hhmm + hhmmss;
        const uint64_t ko[3] = {0, 5, 13}, vo[3] = {0, hhmm.size(), hhmm.size() + hhmmss.size()};
        const uint8_t tags[2] = {IE_TAG_STRING, IE_TAG_STRING};
        const uint32_t ids[2] = {(uint32_t)entry_vals.size(), (uint32_t)entry_vals.size() + 1};
        const ie_status_t st = ie_table_set(e, table, 0, 2, (const uint8_t*)keys.data(), ko, (const uint8_t*)vals.data(), vo, tags, ids);
        if (st == IE_E_OVERFLOW) { pack(); return; }
        check(st);
        packed_hhmm = hhmm;
        packed_hhmmss = hhmmss;
    }

    // interp.rs:122-134: `<dir>/<key>.json5` (parsed, recursive_escape'd) then `<dir>/<key>` (trimmed, escaped)
    bool probe_dir(const std::string& key);

    std::vector<Outcome> resolve(const std::vector<std::string>& templates) {
        std::vector<Outcome> out(templates.size());
        if (templates.empty()) return out;
        if (!table) pack();
        Arena t;
        for (auto& s : templates) t.push(s);
        for (int round = 0; round < 64; ++round) {
            ie_result r;
            check(ie_resolve_batch(e, table, t.data(), t.offs.data(), t.n(), nullptr, &r));
            bool grew = false;
            // First everything out of the engine's result buffers (they are only valid until the next call on the
            // engine) ...
            for (size_t i = 0; i < templates.size(); ++i) {
                Outcome& oc = out[i];
                oc.code = IE_RES_CODE(r.status[i]);
                oc.bytes.assign((const char*)r.out + r.out_offs[i], r.out_lens[i]);
                if (oc.code == IE_RES_TYPED) oc.typed = entry_value(r.aux[i]);
            }
            // ... then the directory probes: escaping a file's value is another call on the same engine
            if (has_dir)
                for (Outcome& oc : out) {
                    if (oc.code != IE_RES_NOT_FOUND) continue;
                    if (probe_dir(oc.bytes)) grew = true;
                    auto de = dir_errors.find(oc.bytes);
                    if (de != dir_errors.end()) oc.io_error = de->second;
                }
            if (!grew) break;
            pack();
        }
        return out;
    }
};

JVal escape_tree(ie_engine* e, int mode, const JVal& v);

bool Session::probe_dir(const std::string& key) {
    if (dir_probed.count(key) || is_arg_key(key)) return false;  // ARG keys never fall back (interp.rs:109-116)
    dir_probed[key] = true;
    if (key.find('\0') != std::string::npos) return false;
    struct stat st;
    const std::string j5 = inserts_dir + "/" + key + ".json5", plain = inserts_dir + "/" + key;
    auto slurp = [](const std::string& path, std::string* o) { std::ifstream f(path, std::ios::binary); if (!f) return false; std::stringstream ss; ss << f.rdbuf(); *o = ss.str(); return true; };
    std::string raw;
    if (::stat(j5.c_str(), &st) == 0) {
        if (!slurp(j5, &raw)) { dir_errors[key] = "cannot read " + j5; return false; }
        try { inserts[key] = escape_tree(e, 1, parse(raw)); } catch (const std::exception& ex) { dir_errors[key] = ex.what(); return false; }
        return true;
    }
    if (::stat(plain.c_str(), &st) == 0) {
        if (!slurp(plain, &raw)) { dir_errors[key] = "cannot read " + plain; return false; }
        size_t a = 0, b = raw.size();
        auto ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
        while (a < b && ws((unsigned char)raw[a])) ++a;
        while (b > a && ws((unsigned char)raw[b - 1])) --b;
        inserts[key] = escape_tree(e, 1, JVal::str(raw.substr(a, b - a)));
        return true;
    }
    return false;
}

JVal outcome_value(const Outcome& oc) {
    if (oc.code == IE_RES_STRING) return JVal::str(oc.bytes);
    if (oc.code == IE_RES_TYPED) return oc.typed;
    if (!oc.io_error.empty()) throw ApiError{9, oc.io_error, oc.bytes};
    throw ApiError{oc.code, message_for(oc.code, oc.bytes), oc.bytes};
}

// ---- recursive_interpolate (interp.rs:179-246) as gather -> one GPU batch -> rebuild -------------
struct Gather {
    std::vector<std::string> templates;
    std::vector<std::string> lookups;  // simple keys of for/serial/parallel_* `tasks` (get_interpdata, :217, :224)
};
bool is_control_cmd(const std::string& c) { return c == "for" || c == "serial" || c == "parallel_wait" || c == "parallel_race"; }
const std::string* cmd_of(const JVal& v) {
    auto it = v.o->find("cmd");
    return (it != v.o->end() && it->second.t == JVal::Str) ? &it->second.s : nullptr;
}
void gather(const JVal& v, Gather& g) {
    if (v.t == JVal::Str) { g.templates.push_back(v.s); return; }
    if (v.t == JVal::Arr) { for (auto& x : *v.a) gather(x, g); return; }
    if (v.t != JVal::Obj) return;
    if (const std::string* cmd = cmd_of(v)) {
        if (*cmd == "goto_map" || *cmd == "replace_map") return;  // :210-212
        if (is_control_cmd(*cmd)) {                              // :213-233
            auto t = v.o->find("tasks");
            if (t != v.o->end()) {
                std::string k;
                if (t->second.t == JVal::Str) { if (simple_insertkey(t->second.s, &k)) g.lookups.push_back(k); }
                else if (t->second.t == JVal::Arr) for (auto& x : *t->second.a) if (x.t == JVal::Str && simple_insertkey(x.s, &k)) g.lookups.push_back(k);
            }
            return;
        }
    }
    for (auto& kv : *v.o) { g.templates.push_back(kv.first); gather(kv.second, g); }
}
struct Rebuild {
    const std::vector<Outcome>& res;
    const std::vector<JVal>& looked;
    const std::vector<std::shared_ptr<ApiError>>& look_errs;  // raised where the traversal reaches the lookup (:217, :224 `?`)
    size_t ti = 0, li = 0;
    JVal lookup_result() {
        if (look_errs[li]) throw *look_errs[li];
        return looked[li++];
    }
    JVal string_result(const std::string& original) {
        const Outcome& oc = res[ti++];
        if (oc.code == IE_RES_STRING) return JVal::str(oc.bytes);
        if (oc.code == IE_RES_TYPED) return oc.typed;
        if (oc.code == IE_RES_PANIC || oc.code == IE_RES_LIMIT) throw ApiError{oc.code, message_for(oc.code, oc.bytes), oc.bytes};
        return JVal::str(original);  // interp.rs:192, :201 — resolver errors are swallowed
    }
    JVal walk(const JVal& v) {
        if (v.t == JVal::Str) return string_result(v.s);
        if (v.t == JVal::Arr) { JArr a; for (auto& x : *v.a) a.push_back(walk(x)); return JVal::arr(std::move(a)); }
        if (v.t != JVal::Obj) return v;
        if (const std::string* cmd = cmd_of(v)) {
            if (*cmd == "goto_map" || *cmd == "replace_map") return v;
            if (is_control_cmd(*cmd)) {
                JObj o = *v.o;
                auto t = o.find("tasks");
                if (t != o.end()) {
                    if (t->second.t == JVal::Str) { if (simple_insertkey(t->second.s, nullptr)) t->second = lookup_result(); }
                    else if (t->second.t == JVal::Arr) {
                        JArr a = *t->second.a;
                        for (auto& x : a) if (x.t == JVal::Str && simple_insertkey(x.s, nullptr)) x = lookup_result();
                        t->second = JVal::arr(std::move(a));
                    }
                }
                return JVal::obj(std::move(o));
            }
        }
        JObj out;
        for (auto& kv : *v.o) {
            JVal nk = string_result(kv.first);
            JVal nv = walk(kv.second);
            out[value_to_string(nk)] = std::move(nv);  // later duplicates win (:241)
        }
        return JVal::obj(std::move(out));
    }
};

// get_interpdata (interp.rs:91-137) for literal keys through the device table
// `errs` (optional): a failed lookup is recorded there instead of thrown, so that the caller can raise it where the
// reference would have reached it (recursive_interpolate evaluates strings and lookups in one traversal order).
std::vector<JVal> lookup_keys(Session& s, const std::vector<std::string>& keys, std::vector<std::shared_ptr<ApiError>>* errs = nullptr) {
    std::vector<JVal> out(keys.size());
    if (errs) errs->assign(keys.size(), nullptr);
    auto failed = [&](size_t i, ApiError er) { if (!errs) throw er; (*errs)[i] = std::make_shared<ApiError>(std::move(er)); };
    if (keys.empty()) return out;
    if (!s.table) s.pack();
    for (int round = 0; round < 64; ++round) {
        Arena a;
        for (auto& k : keys) a.push(k);
        std::vector<int32_t> tag(keys.size());
        std::vector<uint32_t> entry(keys.size());
        check(ie_lookup_batch(s.e, s.table, a.data(), a.offs.data(), a.n(), tag.data(), entry.data()));
        bool grew = false;
        for (size_t i = 0; i < keys.size(); ++i) {
            const std::string& k = keys[i];
            if (tag[i] >= 0) { out[i] = s.entry_value(entry[i]); if (errs) (*errs)[i] = nullptr; continue; }
            if (k.empty()) { failed(i, ApiError{IE_RES_EMPTY_KEY, message_for(IE_RES_EMPTY_KEY, ""), ""}); continue; }
            if (is_arg_key(k)) { failed(i, ApiError{IE_RES_ARG_MISSING, message_for(IE_RES_ARG_MISSING, k), k}); continue; }
            if (s.has_dir && s.probe_dir(k)) { grew = true; continue; }
            auto de = s.dir_errors.find(k);
            if (de != s.dir_errors.end()) { failed(i, ApiError{9, de->second, k}); continue; }
            failed(i, ApiError{IE_RES_NOT_FOUND, message_for(IE_RES_NOT_FOUND, k), k});
        }
        if (!grew) break;
        s.pack();
    }
    return out;
}

// recursive_escape / recursive_unescape (interp.rs:147-177): strings and object keys, one GPU batch
void gather_strings(const JVal& v, std::vector<std::string>& g) {
    if (v.t == JVal::Str) g.push_back(v.s);
    else if (v.t == JVal::Arr) for (auto& x : *v.a) gather_strings(x, g);
    else if (v.t == JVal::Obj) for (auto& kv : *v.o) { g.push_back(kv.first); gather_strings(kv.second, g); }
}
JVal rebuild_strings(const JVal& v, const std::vector<std::string>& r, size_t& i) {
    if (v.t == JVal::Str) return JVal::str(r[i++]);
    if (v.t == JVal::Arr) { JArr a; for (auto& x : *v.a) a.push_back(rebuild_strings(x, r, i)); return JVal::arr(std::move(a)); }
    if (v.t == JVal::Obj) { JObj o; for (auto& kv : *v.o) { std::string k = r[i++]; o[k] = rebuild_strings(kv.second, r, i); } return JVal::obj(std::move(o)); }
    return v;
}
JVal escape_tree(ie_engine* e, int mode, const JVal& v) {
    std::vector<std::string> g;
    gather_strings(v, g);
    if (g.empty()) return v;
    Arena a;
    for (auto& s : g) a.push(s);
    const uint8_t* ob; const uint64_t* oo;
    check(ie_escape_batch(e, mode, a.data(), a.offs.data(), a.n(), &ob, &oo));
    std::vector<std::string> r(g.size());
    for (size_t i = 0; i < g.size(); ++i) r[i].assign((const char*)ob + oo[i], oo[i + 1] - oo[i]);
    size_t i = 0;
    return rebuild_strings(v, r, i);
}

// interp.rs:273-312 — analyzer helper (analyzer.rs:783); a scan, not a resolve: no table involved
void extract_from_str(const std::string& s, JArr& keys) {
    long depth = 0;
    std::string cur;
    bool in_key = false, escaped = false;
    for (char ch : s) {
        if (escaped) { escaped = false; if (in_key) cur += ch; continue; }
        if (ch == '\\') { escaped = true; continue; }
        if (ch == '{') { if (++depth == 1) { in_key = true; cur.clear(); continue; } }
        if (ch == '}') {
            if (depth == 1 && in_key) { keys.push_back(JVal::str(cur)); in_key = false; --depth; continue; }
            if (depth > 0) --depth;
        }
        if (in_key) cur += ch;
    }
}
void extract_keys(const JVal& v, JArr& keys) {
    if (v.t == JVal::Str) extract_from_str(v.s, keys);
    else if (v.t == JVal::Arr) for (auto& x : *v.a) extract_keys(x, keys);
    else if (v.t == JVal::Obj) for (auto& kv : *v.o) { extract_from_str(kv.first, keys); extract_keys(kv.second, keys); }
}

const JVal& arg(const JVal& args, const char* name) {
    static const JVal null_v;
    if (args.t != JVal::Obj) return null_v;
    auto it = args.o->find(name);
    return it == args.o->end() ? null_v : it->second;
}
const std::string& sarg(const JVal& args, const char* name) {
    const JVal& v = arg(args, name);
    if (v.t != JVal::Str) throw std::runtime_error(std::string("missing string argument '") + name + "'");
    return v.s;
}

// ---- the two callers that loop over the resolver (SURVEY.md §8 f): replace_map, goto_map ----------------------
// Every resolve and every wildcard test goes through the GPU batch entry points; what stays on the host is the
// sequencing (the reference evaluates map entries in order and stops at the first match or error) and the
// capture extraction of the one pattern that matched.

ApiError task_error(const std::string& msg) { return ApiError{-2, msg, ""}; }

// runtime.rs:1754-1775: what each '*' swallows under the greedy, leftmost semantics of "^lit(.*)lit...$".
// `ok[j][pos]`: literal j can start at pos and the rest of the pattern still reaches the end of the text.
std::vector<std::string> wildcard_captures(const std::string& p, const std::string& s) {
    std::vector<std::string> lits(1);
    for (char c : p) { if (c == '*') lits.emplace_back(); else lits.back().push_back(c); }
    const size_t m = lits.size() - 1, n = s.size();
    std::vector<std::vector<uint8_t>> ok(m + 1, std::vector<uint8_t>(n + 2, 0)), later(m + 1, std::vector<uint8_t>(n + 2, 0));
    for (size_t j = m + 1; j-- > 0;) {
        const std::string& L = lits[j];
        for (size_t pos = n + 1; pos-- > 0;) {
            bool fits = pos + L.size() <= n && s.compare(pos, L.size(), L) == 0;
            if (fits) fits = j == m ? pos + L.size() == n : later[j + 1][pos + L.size()] != 0;
            ok[j][pos] = fits;
            later[j][pos] = fits || later[j][pos + 1];  // some start >= pos works
        }
    }
    std::vector<std::string> caps;
    if (!ok[0][0] || m == 0) return caps;
    size_t pos = lits[0].size();
    for (size_t j = 1; j <= m; ++j) {
        size_t e = n;
        while (!ok[j][e]) --e;  // the largest start of literal j that still lets the rest match (exists: ok[0][0])
        caps.push_back(s.substr(pos, e - pos));
        pos = e + lits[j].size();
    }
    return caps;
}

// index of the first pattern that matches `text`, -1 if none: one launch (one CTA works on the text, ie_glob.cu)
int first_match(ie_engine* e, const std::string& text, const std::vector<std::string>& patterns) {
    Arena keys, pats;
    keys.push(text);
    for (auto& p : patterns) pats.push(p);
    uint32_t first = 0xFFFFFFFFu;
    check(ie_glob_first_match(e, keys.data(), keys.offs.data(), 1, pats.data(), pats.offs.data(), (uint32_t)patterns.size(), &first));
    return first == 0xFFFFFFFFu ? -1 : (int)first;
}

struct MapEntry { bool is_obj = false, empty = true; std::string key, val; };
std::vector<MapEntry> map_entries(const JVal& maps) {
    std::vector<MapEntry> out;
    for (auto& m : *maps.a) {
        MapEntry en;
        en.is_obj = m.t == JVal::Obj;
        if (en.is_obj && !m.o->empty()) {
            en.empty = false;
            en.key = m.o->begin()->first;  // obj.iter().next(): first key in sorted order
            const JVal& v = m.o->begin()->second;
            en.val = v.t == JVal::Str ? v.s : std::string();  // v.as_str().unwrap_or("")
        }
        out.push_back(en);
    }
    return out;
}

// runtime.rs:1658-1692.  Per iteration: ONE resolve batch (the text and every map key), one first-match sweep,
// and — for the entry that matched — one resolve of its value against inserts + captures.
std::string replace_str(ie_engine* e, const JVal& args, Session& s, std::string text, const std::vector<MapEntry>& maps, bool repeat) {
    std::set<std::string> seen;  // a text that comes back can only repeat its cycle: the reference would not terminate
    for (int guard = 0;; ++guard) {
        if (guard > 10000 || !seen.insert(text).second) throw ApiError{IE_RES_LIMIT, message_for(IE_RES_LIMIT, ""), ""};
        std::vector<std::string> batch{text};
        for (auto& m : maps) if (m.is_obj && !m.empty) batch.push_back(m.key);
        const std::vector<Outcome> res = s.resolve(batch);
        const std::string current = value_to_string(outcome_value(res[0]));
        // entries in order up to the first one the reference would fail on
        std::vector<std::string> keys;
        std::vector<size_t> which;
        const ApiError* pending = nullptr;
        ApiError pending_store{0, "", ""};
        size_t bi = 1;
        for (size_t j = 0; j < maps.size() && !pending; ++j) {
            if (!maps[j].is_obj) { pending_store = task_error("replace_map expects object"); pending = &pending_store; break; }
            if (maps[j].empty) { pending_store = task_error("replace_map entry empty"); pending = &pending_store; break; }
            try { keys.push_back(value_to_string(outcome_value(res[bi++]))); which.push_back(j); }
            catch (const ApiError& er) { pending_store = er; pending = &pending_store; }
        }
        const int f = first_match(e, current, keys);
        if (f < 0 && pending) throw *pending;
        std::string new_text = current;
        if (f >= 0) {
            // runtime.rs:1688-1694: the value is resolved against inserts + {"1": capture 1, ...}.  The captures are set on
            // the session's own table in place and taken back afterwards (set_interpdata / delete_interpdata on the device
            // table) instead of packing inserts + captures as a new snapshot per iteration.
            const std::vector<std::string> caps = wildcard_captures(keys[f], current);
            std::vector<std::pair<std::string, std::pair<bool, JVal>>> saved;
            for (size_t i = 0; i < caps.size(); ++i) {
                const std::string k = std::to_string(i + 1);
                auto it = s.inserts.find(k);
                saved.push_back({k, {it != s.inserts.end(), it != s.inserts.end() ? it->second : JVal()}});
                s.set(k, JVal::str(caps[i]));
            }
            auto restore = [&]() { for (auto& sv : saved) { if (sv.second.first) s.set(sv.first, sv.second.second); else s.del(sv.first); } };
            try { new_text = value_to_string(outcome_value(s.resolve({maps[which[f]].val})[0])); }
            catch (...) { restore(); throw; }
            restore();
        }
        if (!repeat || new_text == text) return new_text;
        text = new_text;
    }
}

// runtime.rs:1733-1752
bool find_null_map_value(Session& s, const JVal& maps, JVal* out) {
    for (auto& m : *maps.a) {
        if (m.t != JVal::Obj) continue;
        for (auto& kv : *m.o) {
            if (kv.first == "NULL") { *out = kv.second; return true; }
            if (kv.first.find('{') != std::string::npos) {
                const Outcome oc = s.resolve({kv.first})[0];
                if (oc.code == IE_RES_PANIC) throw ApiError{oc.code, message_for(oc.code, oc.bytes), oc.bytes};
                if ((oc.code == IE_RES_STRING || oc.code == IE_RES_TYPED) && value_to_string(outcome_value(oc)) == "NULL") { *out = kv.second; return true; }
            }
        }
    }
    return false;
}

// runtime.rs:1649-1731 (every `?` in the match arms leaves the function: errors propagate, the trailing NULL
// handler only ever sees Ok)
JVal replace_map(ie_engine* e, const JVal& args, Session& s, const JVal& item, const std::vector<MapEntry>& maps, bool has_null,
                 const JVal& null_value, bool repeat) {
    if (item.t == JVal::Str) {
        if (has_null && simple_insertkey(item.s, nullptr)) {
            const Outcome oc = s.resolve({item.s})[0];
            if (oc.code == IE_RES_PANIC) throw ApiError{oc.code, message_for(oc.code, oc.bytes), oc.bytes};
            if (oc.code != IE_RES_STRING && oc.code != IE_RES_TYPED) return null_value;
        }
        return JVal::str(replace_str(e, args, s, item.s, maps, repeat));
    }
    if (item.t == JVal::Arr) {
        JArr out;
        for (auto& v : *item.a) out.push_back(replace_map(e, args, s, v, maps, has_null, null_value, repeat));
        return JVal::arr(std::move(out));
    }
    if (item.t == JVal::Obj) {
        JObj out;
        for (auto& kv : *item.o) {
            const std::string nk = replace_str(e, args, s, kv.first, maps, repeat);
            out[nk] = replace_map(e, args, s, kv.second, maps, has_null, null_value, repeat);
        }
        return JVal::obj(std::move(out));
    }
    return item;
}

// runtime.rs:1085-1133: one resolve batch (text, every key, every value), one first-match sweep.
JVal goto_map(ie_engine* e, Session& s, const std::string& text, const JVal& target_maps) {
    const std::vector<MapEntry> maps = map_entries(target_maps);
    std::vector<std::string> batch{text};
    for (auto& m : maps) if (m.is_obj && !m.empty) { batch.push_back(m.key); batch.push_back(m.val); }
    const std::vector<Outcome> res = s.resolve(batch);
    bool interp_error = false;
    std::string value_text;
    if (res[0].code == IE_RES_PANIC) throw ApiError{res[0].code, message_for(res[0].code, res[0].bytes), res[0].bytes};
    try { value_text = value_to_string(outcome_value(res[0])); } catch (const ApiError&) { interp_error = true; value_text = "NULL"; }
    std::vector<std::string> keys, vals;
    std::vector<const Outcome*> val_oc;
    ApiError pending{0, "", ""};
    bool has_pending = false;
    size_t bi = 1;
    for (size_t j = 0; j < maps.size(); ++j) {
        if (!maps[j].is_obj) { pending = task_error("target_maps entry must be object"); has_pending = true; break; }
        if (maps[j].empty) { pending = task_error("target_maps entry empty"); has_pending = true; break; }
        const Outcome& ko = res[bi++];
        const Outcome& vo = res[bi++];
        try { keys.push_back(value_to_string(outcome_value(ko))); } catch (const ApiError& er) { pending = er; has_pending = true; break; }
        if (!interp_error) {  // the value is resolved before the key is tested (:1124-1125)
            try { value_to_string(outcome_value(vo)); } catch (const ApiError& er) { keys.pop_back(); pending = er; has_pending = true; break; }
        }
        val_oc.push_back(&vo);
    }
    int f = -1;
    if (interp_error) { for (size_t j = 0; j < keys.size() && f < 0; ++j) if (keys[j] == "NULL") f = (int)j; }
    else f = first_match(e, value_text, keys);
    if (f < 0) {
        if (has_pending) throw pending;
        if (interp_error) throw task_error("goto_map value could not be resolved but 'NULL' is not a key in target_maps");
        throw task_error("goto_map has no matches for '" + value_text + "'");
    }
    const std::string target = value_to_string(outcome_value(*val_oc[f]));  // interp_error: resolved only now (:1110), errors propagate
    return JVal::obj({{"value", JVal::str(value_text)}, {"target", JVal::str(target)}, {"interpolation_error", JVal::boolean(interp_error)}});
}

// ---- program loader: parser.rs:8-93 -------------------------------------------------------
// Host-only: the on-disk format in front of the path.  `line:N` is spliced behind every `cmd: '<name>'` pair the
// reference's per-line regex (parser.rs:75-77) would match, then the text goes through the JSON5-subset reader.
struct LineScanner {
    const std::string& ln;
    static bool word(unsigned char c) { return std::isalnum(c) || c == '_' || c >= 0x80; }  // \w, ASCII + opaque UTF-8
    static bool space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }           // \s, ASCII
    size_t skip_space(size_t i) const { while (i < ln.size() && space((unsigned char)ln[i])) ++i; return i; }
    // end of the `cmd` key spelled at i (bare word with \b on both sides, or quoted), 0 if none
    size_t key_end(size_t i) const {
        if (ln.compare(i, 3, "cmd") == 0) {
            const bool left = i == 0 || !word((unsigned char)ln[i - 1]);
            const bool right = i + 3 >= ln.size() || !word((unsigned char)ln[i + 3]);
            return left && right ? i + 3 : 0;
        }
        if (ln.compare(i, 5, "\"cmd\"") == 0 || ln.compare(i, 5, "'cmd'") == 0) return i + 5;
        return 0;
    }
    // end (one past the closing quote) of the string literal that starts at i, 0 if it does not close on this line
    size_t string_end(size_t i) const {
        if (i >= ln.size() || (ln[i] != '"' && ln[i] != '\'')) return 0;
        const char q = ln[i];
        for (size_t v = i + 1; v < ln.size(); ++v) {
            if (ln[v] == '\\') { if (++v >= ln.size()) return 0; continue; }
            if (ln[v] == q) return v + 1;
        }
        return 0;
    }
};
std::string add_line_numbers(const std::string& text) {
    std::string out;
    size_t begin = 0, number = 0;
    while (begin < text.size()) {  // str::lines()
        size_t end = text.find('\n', begin);
        const bool had_nl = end != std::string::npos;
        if (!had_nl) end = text.size();
        std::string ln = text.substr(begin, end - begin);
        begin = had_nl ? end + 1 : end;
        if (had_nl && !ln.empty() && ln.back() == '\r') ln.pop_back();
        ++number;
        const LineScanner sc{ln};
        size_t from = 0;
        for (size_t i = 0; i < ln.size();) {
            const size_t k = sc.key_end(i);
            size_t c = k ? sc.skip_space(k) : 0;
            if (k && c < ln.size() && ln[c] == ':') {
                const size_t v0 = sc.skip_space(c + 1), v1 = sc.string_end(v0);
                const size_t t = v1 ? sc.skip_space(v1) : 0;
                if (v1 && t < ln.size() && (ln[t] == ',' || ln[t] == '}')) {
                    out.append(ln, from, i - from).append(ln, i, k - i).append(":").append(ln, v0, v1 - v0);
                    out.append(", line:").append(std::to_string(number)).append(ln, v1, t + 1 - v1);
                    i = from = t + 1;
                    continue;
                }
            }
            ++i;
        }
        out.append(ln, from, std::string::npos).push_back('\n');
    }
    return out;
}
JVal load_program(const std::string& raw) {
    JVal root = parse(add_line_numbers(raw));
    if (root.t != JVal::Obj) throw task_error("Program root must be an object");
    JObj& obj = *root.o;
    if (!obj.count("named_tasks") && obj.count("tasks")) { obj["named_tasks"] = obj["tasks"]; obj.erase("tasks"); }  // parser.rs:17-20
    auto object_field = [&](const char* k) -> JVal {
        auto it = obj.find(k);
        if (it == obj.end() || it->second.t != JVal::Obj) throw task_error(std::string("Program missing '") + k + "' object");
        return it->second;
    };
    auto check_task = [](const JVal& v) {
        if (v.t == JVal::Obj) return;
        std::string j;
        dump(v, j);  // the reference formats the value with {:?}; compact JSON here (documented deviation)
        throw task_error("Task must be an object, got " + j);
    };
    JObj out;
    out["default_state"] = object_field("default_state");
    auto ord = obj.find("order");
    if (ord == obj.end() || ord->second.t != JVal::Arr) throw task_error("Program missing 'order' array");
    for (auto& t : *ord->second.a) check_task(t);
    out["order"] = ord->second;
    const JVal named = object_field("named_tasks");
    for (auto& kv : *named.o) check_task(kv.second);
    out["named_tasks"] = named;
    out["save_states"] = object_field("save_states");
    auto ca = obj.find("completion_args");
    out["completion_args"] = (ca != obj.end() && ca->second.t == JVal::Obj) ? ca->second : JVal::obj();
    return JVal::obj(std::move(out));
}

// ---- snapshots that outlive a call ------------------------------------------------------------------------------------
// The reference keeps ONE inserts map per run, mutates it between tasks (set_interpdata, 17 call sites in runtime.rs) and
// resolves against it task after task.  A caller that follows that pattern creates a snapshot once ("snapshot_create"),
// mirrors its set_interpdata / delete_interpdata calls ("snapshot_set" / "snapshot_delete": the packed table is patched in
// place) and passes {"snapshot": id} instead of {"inserts": {...}} to every other function: no per-call serialisation,
// packing or upload of the state.
std::mutex g_snap_mu;
std::map<std::pair<ie_engine*, uint64_t>, std::unique_ptr<Session>> g_snaps;
uint64_t g_snap_next = 1;

struct Hold {  // the session of one call: a registered snapshot (clock refreshed) or a temporary built from "inserts"
    std::unique_ptr<Session> own;
    Session* s = nullptr;
    Hold(ie_engine* e, const JVal& args) {
        const JVal* id = nullptr;
        if (args.t == JVal::Obj) { auto it = args.o->find("snapshot"); if (it != args.o->end()) id = &it->second; }
        if (!id) { own.reset(new Session(e, args)); s = own.get(); return; }
        std::lock_guard<std::mutex> lk(g_snap_mu);
        auto it = g_snaps.find({e, (uint64_t)std::strtoull(id->s.c_str(), nullptr, 10)});
        if (id->t != JVal::Num || it == g_snaps.end()) throw std::runtime_error("unknown snapshot");
        s = it->second.get();
        std::string hm, hs;
        Session::read_clock(args, &hm, &hs);
        s->refresh_clock(hm, hs);
        auto d = args.o->find("inserts_dir");
        s->has_dir = d != args.o->end() && d->second.t == JVal::Str;
        if (s->has_dir && s->inserts_dir != d->second.s) { s->inserts_dir = d->second.s; s->dir_probed.clear(); s->dir_errors.clear(); }
    }
};

JVal dispatch(ie_engine* e, const JVal& args) {
    const std::string& fn = sarg(args, "fn");
    if (fn == "snapshot_create") {
        std::unique_ptr<Session> ns(new Session(e, args));
        ns->pack();
        std::lock_guard<std::mutex> lk(g_snap_mu);
        const uint64_t id = g_snap_next++;
        g_snaps[{e, id}] = std::move(ns);
        return JVal::num(std::to_string(id));
    }
    if (fn == "snapshot_set" || fn == "snapshot_delete" || fn == "snapshot_free" || fn == "snapshot_inserts") {
        const JVal& idv = arg(args, "snapshot");
        std::lock_guard<std::mutex> lk(g_snap_mu);
        auto it = g_snaps.find({e, (uint64_t)std::strtoull(idv.s.c_str(), nullptr, 10)});
        if (idv.t != JVal::Num || it == g_snaps.end()) throw std::runtime_error("unknown snapshot");
        if (fn == "snapshot_free") { g_snaps.erase(it); return JVal(); }
        if (fn == "snapshot_inserts") return JVal::obj(JObj(it->second->inserts));  // the map as the host mirror holds it
        if (fn == "snapshot_set") it->second->set(sarg(args, "key"), arg(args, "value"));  // interp.rs:139
        else it->second->del(sarg(args, "key"));                                           // interp.rs:143
        return JVal();
    }
    if (fn == "interpolate_inserts") {  // interp.rs:31
        Hold h(e, args); Session& s = *h.s;
        return outcome_value(s.resolve({sarg(args, "content")})[0]);
    }
    if (fn == "interpolate_many") {  // batch form: [{ok|err}] per template, one launch
        Hold h(e, args); Session& s = *h.s;
        std::vector<std::string> ts;
        for (auto& x : *arg(args, "contents").a) ts.push_back(x.s);
        JArr out;
        for (auto& oc : s.resolve(ts)) {
            JObj r;
            try { r["ok"] = outcome_value(oc); }
            catch (const ApiError& er) { r["err"] = JVal::obj({{"code", JVal::num(std::to_string(er.code))}, {"message", JVal::str(er.message)}, {"payload", JVal::str(er.payload)}}); }
            out.push_back(JVal::obj(std::move(r)));
        }
        return JVal::arr(std::move(out));
    }
    if (fn == "get_simple_insertkey") {  // interp.rs:11
        std::string k;
        return simple_insertkey(sarg(args, "content"), &k) ? JVal::str(k) : JVal();
    }
    if (fn == "get_interpdata") {  // interp.rs:91
        Hold h(e, args); Session& s = *h.s;
        return lookup_keys(s, {sarg(args, "key")})[0];
    }
    if (fn == "recursive_interpolate") {  // interp.rs:179
        Hold h(e, args); Session& s = *h.s;
        const JVal& v = arg(args, "value");
        Gather g;
        gather(v, g);
        std::vector<std::shared_ptr<ApiError>> look_errs;
        const std::vector<JVal> looked = lookup_keys(s, g.lookups, &look_errs);
        const std::vector<Outcome> res = s.resolve(g.templates);
        Rebuild rb{res, looked, look_errs};
        return rb.walk(v);
    }
    if (fn == "recursive_escape") return escape_tree(e, 1, arg(args, "value"));      // interp.rs:163
    if (fn == "recursive_unescape") return escape_tree(e, 0, arg(args, "value"));    // interp.rs:147
    if (fn == "value_to_string") return JVal::str(value_to_string(arg(args, "value")));  // interp.rs:314
    if (fn == "extract_insert_keys") { JArr k; extract_keys(arg(args, "value"), k); return JVal::arr(std::move(k)); }  // interp.rs:248
    if (fn == "wildcard_match" || fn == "delete" || fn == "delete_except") {
        Arena keys, pats;
        JObj ins;
        if (fn == "wildcard_match") {  // runtime.rs:1633
            keys.push(sarg(args, "text"));
            pats.push(sarg(args, "pattern"));
        } else {  // runtime.rs:1198-1239: keys in sorted order; wildcards through value_to_string
            const JVal& iv = arg(args, "inserts");
            if (iv.t == JVal::Obj) ins = *iv.o;
            for (auto& kv : ins) keys.push(kv.first);
            const JVal& w = arg(args, "wildcards");
            if (w.t == JVal::Arr) for (auto& x : *w.a) pats.push(value_to_string(x));
        }
        JArr deleted;
        // more than IE_MAX_PATTERNS wildcards: sweep in groups; a key is matched when any group matches it
        const uint64_t n = keys.n(), words = (n + 31) / 32;
        std::vector<uint32_t> any(words, 0), mask(words + 1, 0);
        for (uint64_t p0 = 0; p0 < pats.n(); p0 += IE_MAX_PATTERNS) {
            const uint32_t np = (uint32_t)std::min<uint64_t>(IE_MAX_PATTERNS, pats.n() - p0);
            std::vector<uint64_t> po(np + 1);
            for (uint32_t i = 0; i <= np; ++i) po[i] = pats.offs[p0 + i] - pats.offs[p0];
            uint64_t nd = 0;
            check(ie_glob_sweep(e, keys.data(), keys.offs.data(), n, pats.data() + pats.offs[p0], po.data(), np, 0, mask.data(), &nd));
            for (uint64_t w = 0; w < words; ++w) any[w] |= mask[w];
        }
        if (fn == "wildcard_match") return JVal::boolean(n && (any[0] & 1u));
        const bool except = fn == "delete_except";
        uint64_t k = 0;
        std::vector<std::string> doomed;
        for (auto& kv : ins) { const bool m = (any[k >> 5] >> (k & 31)) & 1u; if (m != except) doomed.push_back(kv.first); ++k; }
        for (auto& d : doomed) { ins.erase(d); deleted.push_back(JVal::str(d)); }
        return JVal::obj({{"deleted", JVal::arr(std::move(deleted))}, {"inserts", JVal::obj(std::move(ins))}});
    }
    if (fn == "interpolation_trace") {  // what recursive_interpolate(value) sends to the resolver, in order: the batch of a replayed program
        Gather g;
        gather(arg(args, "value"), g);
        JArr t, l;
        for (auto& x : g.templates) t.push_back(JVal::str(x));
        for (auto& x : g.lookups) l.push_back(JVal::str(x));
        JObj o;
        o["templates"] = JVal::arr(std::move(t));
        o["lookups"] = JVal::arr(std::move(l));
        return JVal::obj(std::move(o));
    }
    if (fn == "add_line_numbers") return JVal::str(add_line_numbers(sarg(args, "text")));  // parser.rs:74
    if (fn == "load_program") return load_program(sarg(args, "text"));                      // parser.rs:8
    if (fn == "wildcard_captures") {  // runtime.rs:1754
        const std::string &p = sarg(args, "pattern"), &t = sarg(args, "text");
        JArr out;
        if (first_match(e, t, {p}) == 0) for (auto& c : wildcard_captures(p, t)) out.push_back(JVal::str(c));
        return JVal::arr(std::move(out));
    }
    if (fn == "replace_map") {  // runtime.rs:1146-1169, 1649
        const JVal& maps = arg(args, "wildcard_maps");
        if (maps.t != JVal::Arr) throw task_error("replace_map.wildcard_maps must be array");
        Hold h(e, args); Session& s = *h.s;
        JVal null_value;
        const bool has_null = find_null_map_value(s, maps, &null_value);
        const JVal& rep = arg(args, "repeat_until_done");
        return replace_map(e, args, s, arg(args, "item"), map_entries(maps), has_null, null_value, rep.t == JVal::Bool && rep.b);
    }
    if (fn == "goto_map") {  // runtime.rs:1085-1133
        const JVal& maps = arg(args, "target_maps");
        if (maps.t != JVal::Arr) throw task_error("goto_map.target_maps must be array");
        Hold h(e, args); Session& s = *h.s;
        return goto_map(e, s, sarg(args, "text"), maps);
    }
    throw std::runtime_error("unknown fn '" + fn + "'");
}

}  // namespace

void drop_engine(ie_engine* e) {  // ie_engine_destroy: the engine's registered snapshots (and their tables) go first
    std::lock_guard<std::mutex> lk(g_snap_mu);
    for (auto it = g_snaps.begin(); it != g_snaps.end();) it = it->first.first == e ? g_snaps.erase(it) : std::next(it);
}

ie_status_t call_json(ie_engine* e, const std::string& args_json, std::string* out_json, std::string* why) {
    JObj res;
    try {
        const JVal args = parse(args_json);
        res["ok"] = dispatch(e, args);
    } catch (const ApiError& er) {
        res["err"] = JVal::obj({{"code", JVal::num(std::to_string(er.code))}, {"message", JVal::str(er.message)}, {"payload", JVal::str(er.payload)}});
    } catch (const CallFailure& cf) {
        *why = cf.why;
        return cf.st;
    } catch (const std::exception& ex) {
        res["err"] = JVal::obj({{"code", JVal::num("-1")}, {"message", JVal::str(ex.what())}, {"payload", JVal::str("")}});
    }
    out_json->clear();
    dump(JVal::obj(std::move(res)), *out_json);
    return IE_OK;
}

}  // namespace ie_host
