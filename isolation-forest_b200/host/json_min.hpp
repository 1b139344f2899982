// Minimal JSON value (parse + compact render) for the model metadata line
// (IF/core/IsolationForestModelReadWriteUtils.scala:97-187 uses json4s for the same purpose).
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ifb200 {
namespace json {

struct Value {
    enum Kind { Null, Bool, Int, Double, String, Array, Object } kind = Null;
    bool b = false;
    long long i = 0;
    double d = 0.0;
    std::string s;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;  // insertion order preserved

    const Value *get(const std::string &k) const {
        for (auto &kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    bool isNumber() const { return kind == Int || kind == Double; }
    double asDouble() const { return kind == Int ? (double)i : d; }
    long long asInt() const { return kind == Int ? i : (long long)d; }
};

// Shortest decimal string that round-trips, rendered the way json4s/Java print doubles ("256.0", "1.0E-4").
inline std::string javaDouble(double x) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    int prec = 1;
    for (; prec <= 17; prec++) {
        snprintf(buf, sizeof buf, "%.*e", prec - 1, x);
        if (strtod(buf, nullptr) == x) break;
    }
    // buf = [-]d[.ddd]e[+-]XX
    std::string m(buf);
    size_t epos = m.find('e');
    int exp10 = atoi(m.c_str() + epos + 1);
    std::string mant = m.substr(0, epos);
    bool neg = mant[0] == '-';
    if (neg) mant = mant.substr(1);
    std::string digits;
    for (char c : mant)
        if (c != '.') digits += c;
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    std::string out;
    const double ax = std::fabs(x);
    if (ax >= 1e-3 && ax < 1e7) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)exp10 + 1));
            while ((int)ip.size() < exp10 + 1) ip += '0';
            std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr(exp10 + 1) : "0";
            out = ip + "." + fp;
        } else {
            out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
        }
    } else {
        out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
    }
    return neg ? "-" + out : out;
}

inline std::string javaFloat(float x) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    int prec = 1;
    for (; prec <= 9; prec++) {
        snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)x);
        if (strtof(buf, nullptr) == x) break;
    }
    std::string m(buf);
    size_t epos = m.find('e');
    int exp10 = atoi(m.c_str() + epos + 1);
    std::string mant = m.substr(0, epos);
    bool neg = mant[0] == '-';
    if (neg) mant = mant.substr(1);
    std::string digits;
    for (char c : mant)
        if (c != '.') digits += c;
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    std::string out;
    const float ax = std::fabs(x);
    if (ax >= 1e-3f && ax < 1e7f) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)exp10 + 1));
            while ((int)ip.size() < exp10 + 1) ip += '0';
            std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr(exp10 + 1) : "0";
            out = ip + "." + fp;
        } else {
            out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
        }
    } else {
        out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
    }
    return neg ? "-" + out : out;
}

inline void escape(const std::string &s, std::string &out) {
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    char b[8];
                    snprintf(b, sizeof b, "\\u%04x", c);
                    out += b;
                } else {
                    out += (char)c;
                }
        }
    }
    out += '"';
}

inline void render(const Value &v, std::string &out) {
    switch (v.kind) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Int: out += std::to_string(v.i); break;
        case Value::Double: out += javaDouble(v.d); break;
        case Value::String: escape(v.s, out); break;
        case Value::Array:
            out += '[';
            for (size_t i = 0; i < v.arr.size(); i++) {
                if (i) out += ',';
                render(v.arr[i], out);
            }
            out += ']';
            break;
        case Value::Object:
            out += '{';
            for (size_t i = 0; i < v.obj.size(); i++) {
                if (i) out += ',';
                escape(v.obj[i].first, out);
                out += ':';
                render(v.obj[i].second, out);
            }
            out += '}';
            break;
    }
}
inline std::string render(const Value &v) {
    std::string s;
    render(v, s);
    return s;
}

class Parser {
   public:
    explicit Parser(const std::string &t) : t_(t) {}
    Value parse() {
        Value v = value();
        ws();
        if (p_ != t_.size()) fail("trailing characters");
        return v;
    }

   private:
    const std::string &t_;
    size_t p_ = 0;
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("JSON parse error: ") + m); }
    void ws() {
        while (p_ < t_.size() && strchr(" \t\r\n", t_[p_])) p_++;
    }
    Value value() {
        ws();
        if (p_ >= t_.size()) fail("unexpected end");
        char c = t_[p_];
        Value v;
        if (c == '{') {
            v.kind = Value::Object;
            p_++;
            ws();
            if (t_[p_] == '}') { p_++; return v; }
            while (true) {
                ws();
                Value k = str();
                ws();
                if (t_[p_++] != ':') fail("expected ':'");
                v.obj.emplace_back(k.s, value());
                ws();
                if (t_[p_] == ',') { p_++; continue; }
                if (t_[p_] == '}') { p_++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = Value::Array;
            p_++;
            ws();
            if (t_[p_] == ']') { p_++; return v; }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (t_[p_] == ',') { p_++; continue; }
                if (t_[p_] == ']') { p_++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v = str();
        } else if (!t_.compare(p_, 4, "true")) {
            v.kind = Value::Bool; v.b = true; p_ += 4;
        } else if (!t_.compare(p_, 5, "false")) {
            v.kind = Value::Bool; v.b = false; p_ += 5;
        } else if (!t_.compare(p_, 4, "null")) {
            p_ += 4;
        } else {
            size_t s = p_;
            bool isd = false;
            while (p_ < t_.size() && strchr("+-0123456789.eE", t_[p_])) {
                if (strchr(".eE", t_[p_])) isd = true;
                p_++;
            }
            if (s == p_) fail("unexpected character");
            std::string n = t_.substr(s, p_ - s);
            if (isd) { v.kind = Value::Double; v.d = strtod(n.c_str(), nullptr); }
            else { v.kind = Value::Int; v.i = strtoll(n.c_str(), nullptr, 10); }
        }
        return v;
    }
    Value str() {
        if (t_[p_] != '"') fail("expected string");
        p_++;
        Value v;
        v.kind = Value::String;
        while (p_ < t_.size() && t_[p_] != '"') {
            char c = t_[p_++];
            if (c == '\\') {
                char e = t_[p_++];
                switch (e) {
                    case 'n': v.s += '\n'; break;
                    case 't': v.s += '\t'; break;
                    case 'r': v.s += '\r'; break;
                    case 'b': v.s += '\b'; break;
                    case 'f': v.s += '\f'; break;
                    case 'u': {
                        unsigned cp = (unsigned)strtoul(t_.substr(p_, 4).c_str(), nullptr, 16);
                        p_ += 4;
                        if (cp < 0x80) v.s += (char)cp;
                        else if (cp < 0x800) { v.s += (char)(0xC0 | (cp >> 6)); v.s += (char)(0x80 | (cp & 0x3F)); }
                        else { v.s += (char)(0xE0 | (cp >> 12)); v.s += (char)(0x80 | ((cp >> 6) & 0x3F)); v.s += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: v.s += e;
                }
            } else {
                v.s += c;
            }
        }
        if (p_ >= t_.size()) fail("unterminated string");
        p_++;
        return v;
    }
};

inline Value parse(const std::string &t) { return Parser(t).parse(); }
inline Value mkInt(long long i) { Value v; v.kind = Value::Int; v.i = i; return v; }
inline Value mkDouble(double d) { Value v; v.kind = Value::Double; v.d = d; return v; }
inline Value mkBool(bool b) { Value v; v.kind = Value::Bool; v.b = b; return v; }
inline Value mkString(const std::string &s) { Value v; v.kind = Value::String; v.s = s; return v; }
inline Value mkObject() { Value v; v.kind = Value::Object; return v; }
inline Value mkArray() { Value v; v.kind = Value::Array; return v; }

}  // namespace json
}  // namespace ifb200
