// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this code.  The product (interpolation_engine_b200/) never includes it.
//
// A tiny JSON value type standing in for serde_json::Value as the reference uses it
// (rust-project/Cargo.toml:19: serde_json without `preserve_order`, so Map = BTreeMap,
// i.e. object keys iterate in sorted byte order).  Numbers keep their integer/float
// identity the way serde_json::Number does (u64 / i64 / f64).
//
// Parity note: serde_json is not vendored under /root/reference.
// Integer rendering is exact; f64 rendering restates the published ryu "pretty" layout
// (shortest round-trip digits; fixed notation for exponents in [-5,16), always a ".0" on
// integral floats) and is PARITY UNPINNED beyond what the reference's examples exercise.
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct Value;
using Array = std::vector<Value>;
using Object = std::map<std::string, Value>;  // sorted, like BTreeMap

struct Value {
    enum Kind { Null, Bool, Int, UInt, Float, String, Arr, Obj } kind = Null;
    bool b = false;
    int64_t i = 0;
    uint64_t u = 0;
    double f = 0.0;
    std::string s;
    std::shared_ptr<Array> a;
    std::shared_ptr<Object> o;
    // Batch-level inputs arrive pre-rendered (tag + value_to_string text); `raw` marks such a
    // value: `kind` is its logical type and `s` the rendered text (see oracle_capi.cpp).
    bool raw = false;

    static Value null() { return Value(); }
    static Value boolean(bool x) { Value v; v.kind = Bool; v.b = x; return v; }
    static Value integer(int64_t x) { Value v; v.kind = Int; v.i = x; return v; }
    static Value uinteger(uint64_t x) { Value v; v.kind = UInt; v.u = x; return v; }
    static Value real(double x) { Value v; v.kind = Float; v.f = x; return v; }
    static Value string(std::string x) { Value v; v.kind = String; v.s = std::move(x); return v; }
    static Value array(Array x = {}) { Value v; v.kind = Arr; v.a = std::make_shared<Array>(std::move(x)); return v; }
    static Value object(Object x = {}) { Value v; v.kind = Obj; v.o = std::make_shared<Object>(std::move(x)); return v; }
    static Value rendered(Kind logical, std::string text) { Value v; v.kind = logical; v.raw = true; v.s = std::move(text); return v; }

    bool is_string() const { return kind == String; }
    bool is_number() const { return kind == Int || kind == UInt || kind == Float; }
    // serde_json::Value::clone is deep; arrays/objects here are shared_ptr, so make the deep
    // copy explicit where the reference clones (interp.rs:111,119) to keep its cost shape.
    Value deep_clone() const {
        Value v = *this;
        if (raw) return v;
        if (kind == Arr) { v.a = std::make_shared<Array>(); for (auto& e : *a) v.a->push_back(e.deep_clone()); }
        if (kind == Obj) { v.o = std::make_shared<Object>(); for (auto& kv : *o) (*v.o)[kv.first] = kv.second.deep_clone(); }
        return v;
    }
};

// ---- number rendering (serde_json::Number::to_string) -------------------------------
inline std::string f64_to_string(double x) {
    if (!std::isfinite(x)) return "null";  // serde_json cannot hold non-finite numbers
    char digs[64];
    // shortest round-trip digits in scientific form: d.ddddde[+-]XX
    auto r = std::to_chars(digs, digs + sizeof digs, x, std::chars_format::scientific);
    std::string sci(digs, r.ptr);
    bool neg = false;
    size_t p = 0;
    if (sci[0] == '-') { neg = true; p = 1; }
    size_t epos = sci.find('e');
    std::string mant = sci.substr(p, epos - p);
    int exp10 = std::atoi(sci.c_str() + epos + 1);
    std::string d;
    for (char c : mant) if (c != '.') d.push_back(c);
    while (d.size() > 1 && d.back() == '0') d.pop_back();
    int len = (int)d.size();
    int k = exp10 - (len - 1);  // value = d * 10^k
    int kk = len + k;           // position of the decimal point
    std::string out;
    if (neg) out.push_back('-');
    if (x == 0.0) { out += "0.0"; return out; }
    if (0 <= k && kk <= 16) {
        out += d; out.append((size_t)k, '0'); out += ".0";
    } else if (0 < kk && kk <= 16) {
        out += d.substr(0, (size_t)kk); out.push_back('.'); out += d.substr((size_t)kk);
    } else if (-5 < kk && kk <= 0) {
        out += "0."; out.append((size_t)(-kk), '0'); out += d;
    } else if (len == 1) {
        out += d; out.push_back('e'); out += std::to_string(kk - 1);
    } else {
        out += d.substr(0, 1); out.push_back('.'); out += d.substr(1);
        out.push_back('e'); out += std::to_string(kk - 1);
    }
    return out;
}

inline std::string number_to_string(const Value& v) {
    if (v.raw) return v.s;
    switch (v.kind) {
        case Value::Int: return std::to_string(v.i);
        case Value::UInt: return std::to_string(v.u);
        case Value::Float: return f64_to_string(v.f);
        default: return "";
    }
}

// ---- compact JSON writer (serde_json::to_string) --------------------------------------
inline void json_escape_to(const std::string& s, std::string& out) {
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            default:
                if (c < 0x20) { char buf[8]; std::snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
                else out.push_back((char)c);
        }
    }
    out.push_back('"');
}

inline void to_json(const Value& v, std::string& out) {
    switch (v.kind) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Int: case Value::UInt: case Value::Float: out += number_to_string(v); break;
        case Value::String: json_escape_to(v.s, out); break;
        case Value::Arr: {
            out.push_back('[');
            bool first = true;
            for (auto& e : *v.a) { if (!first) out.push_back(','); first = false; to_json(e, out); }
            out.push_back(']');
            break;
        }
        case Value::Obj: {
            out.push_back('{');
            bool first = true;
            for (auto& kv : *v.o) {
                if (!first) out.push_back(',');
                first = false;
                json_escape_to(kv.first, out); out.push_back(':'); to_json(kv.second, out);
            }
            out.push_back('}');
            break;
        }
    }
}
inline std::string to_json(const Value& v) { std::string s; to_json(v, s); return s; }

// ---- JSON / JSON5-subset reader ---------------------------------------------------------
// Accepts strict JSON plus the JSON5 features the reference's example programs use:
// // and /* */ comments, single-quoted strings, unquoted identifier keys, trailing commas,
// leading '+', and line continuations inside strings.
struct Parser {
    const char* p; const char* e;
    explicit Parser(const std::string& s) : p(s.data()), e(s.data() + s.size()) {}
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json parse: ") + m); }
    void ws() {
        for (;;) {
            while (p < e && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
            if (p + 1 < e && p[0] == '/' && p[1] == '/') { while (p < e && *p != '\n') ++p; continue; }
            if (p + 1 < e && p[0] == '/' && p[1] == '*') {
                p += 2; while (p + 1 < e && !(p[0] == '*' && p[1] == '/')) ++p; p += 2; continue;
            }
            break;
        }
    }
    static void put_utf8(uint32_t cp, std::string& out) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    uint32_t hex4() {
        if (e - p < 4) fail("short \\u");
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            char c = *p++; v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else fail("bad hex");
        }
        return v;
    }
    std::string str() {
        char q = *p++;
        std::string out;
        while (p < e && *p != q) {
            char c = *p++;
            if (c != '\\') { out.push_back(c); continue; }
            if (p >= e) fail("dangling escape");
            char n = *p++;
            switch (n) {
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case '0': out.push_back('\0'); break;
                case '\n': break;
                case 'u': {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2; uint32_t lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    put_utf8(cp, out);
                    break;
                }
                default: out.push_back(n);
            }
        }
        if (p >= e) fail("unterminated string");
        ++p;
        return out;
    }
    Value number() {
        const char* s0 = p;
        if (p < e && (*p == '+' || *p == '-')) ++p;
        bool is_float = false;
        while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
            if (*p == '.' || *p == 'e' || *p == 'E') is_float = true;
            ++p;
        }
        std::string t(s0, p);
        if (!t.empty() && t[0] == '+') t.erase(0, 1);
        if (t.empty()) fail("bad number");
        if (!is_float) {
            if (t[0] == '-') { int64_t v = 0; auto r = std::from_chars(t.data(), t.data() + t.size(), v); if (r.ec == std::errc() && r.ptr == t.data() + t.size()) return Value::integer(v); }
            else { uint64_t v = 0; auto r = std::from_chars(t.data(), t.data() + t.size(), v); if (r.ec == std::errc() && r.ptr == t.data() + t.size()) { if (v <= (uint64_t)INT64_MAX) return Value::integer((int64_t)v); return Value::uinteger(v); } }
        }
        return Value::real(std::strtod(t.c_str(), nullptr));
    }
    Value value() {
        ws();
        if (p >= e) fail("unexpected end");
        char c = *p;
        if (c == '{') {
            ++p; Object o;
            for (;;) {
                ws();
                if (p < e && *p == '}') { ++p; break; }
                std::string k;
                if (*p == '"' || *p == '\'') k = str();
                else { const char* s0 = p; while (p < e && (std::isalnum((unsigned char)*p) || *p == '_' || *p == '$')) ++p; if (p == s0) fail("bad key"); k.assign(s0, p); }
                ws();
                if (p >= e || *p != ':') fail("expected ':'");
                ++p;
                Value v = value();
                o[k] = std::move(v);  // later duplicates win
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; break; }
                fail("expected ',' or '}'");
            }
            return Value::object(std::move(o));
        }
        if (c == '[') {
            ++p; Array a;
            for (;;) {
                ws();
                if (p < e && *p == ']') { ++p; break; }
                a.push_back(value());
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; break; }
                fail("expected ',' or ']'");
            }
            return Value::array(std::move(a));
        }
        if (c == '"' || c == '\'') return Value::string(str());
        if (e - p >= 4 && !std::strncmp(p, "true", 4)) { p += 4; return Value::boolean(true); }
        if (e - p >= 5 && !std::strncmp(p, "false", 5)) { p += 5; return Value::boolean(false); }
        if (e - p >= 4 && !std::strncmp(p, "null", 4)) { p += 4; return Value::null(); }
        return number();
    }
};

inline Value parse_json(const std::string& text) {
    Parser ps(text);
    Value v = ps.value();
    ps.ws();
    if (ps.p != ps.e) ps.fail("trailing characters");
    return v;
}

}  // namespace orc
