// mini_json.hpp — a small strict JSON reader (RFC 8259) used on the HOST side only.
//
// Shared by the product's table compiler (cordum_b200/csrc) and by the oracle
// (oracle/).  It carries no policy semantics: it only turns bytes into a tree.
// Object members keep their textual order and duplicates (callers decide the
// duplicate rule; Go's encoding/json is "last one wins").
#pragma once
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace mjson {

enum class Kind : uint8_t { Null, Bool, Number, String, Array, Object };

struct Value {
  Kind kind = Kind::Null;
  bool b = false;
  bool is_int = false;   // number had no fraction/exponent and fits int64
  int64_t i = 0;
  double d = 0.0;
  std::string s;                                      // String
  std::vector<Value> arr;                             // Array
  std::vector<std::pair<std::string, Value>> obj;     // Object (ordered, dups kept)

  bool is_null() const { return kind == Kind::Null; }
  bool is_obj() const { return kind == Kind::Object; }
  bool is_arr() const { return kind == Kind::Array; }
  bool is_str() const { return kind == Kind::String; }
  bool is_num() const { return kind == Kind::Number; }
  bool is_bool() const { return kind == Kind::Bool; }

  // Last member with exactly this key (Go map semantics: last wins), or nullptr.
  const Value* get(std::string_view key) const {
    if (kind != Kind::Object) return nullptr;
    const Value* out = nullptr;
    for (auto& kv : obj)
      if (kv.first == key) out = &kv.second;
    return out;
  }
};

class Parser {
 public:
  Parser(const char* p, size_t n) : p_(p), e_(p + n) {}

  // Parses exactly one JSON value followed only by whitespace.
  bool parse(Value& out, std::string* err = nullptr) {
    ws();
    if (!value(out, 0)) { if (err) *err = err_; return false; }
    ws();
    if (p_ != e_) { if (err) *err = "trailing characters after JSON value"; return false; }
    return true;
  }

 private:
  const char* p_;
  const char* e_;
  std::string err_;

  bool fail(const char* m) { if (err_.empty()) err_ = m; return false; }
  void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }

  bool lit(const char* w) {
    size_t n = std::strlen(w);
    if (size_t(e_ - p_) < n || std::memcmp(p_, w, n) != 0) return fail("invalid literal");
    p_ += n;
    return true;
  }

  static void put_utf8(std::string& s, uint32_t c) {
    if (c < 0x80) s.push_back(char(c));
    else if (c < 0x800) { s.push_back(char(0xC0 | (c >> 6))); s.push_back(char(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) {
      s.push_back(char(0xE0 | (c >> 12))); s.push_back(char(0x80 | ((c >> 6) & 0x3F)));
      s.push_back(char(0x80 | (c & 0x3F)));
    } else {
      s.push_back(char(0xF0 | (c >> 18))); s.push_back(char(0x80 | ((c >> 12) & 0x3F)));
      s.push_back(char(0x80 | ((c >> 6) & 0x3F))); s.push_back(char(0x80 | (c & 0x3F)));
    }
  }

  // length of the valid UTF-8 sequence starting at p (2..4), 0 if p[0] does not start one
  static size_t utf8_len(const char* q, size_t n) {
    const unsigned char* p = (const unsigned char*)q;
    unsigned c = p[0];
    if (c < 0xC2 || c > 0xF4) return 0;
    size_t need = c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : 1;
    if (n < need + 1) return 0;
    unsigned lo = 0x80, hi = 0xBF;
    if (c == 0xE0) lo = 0xA0; else if (c == 0xED) hi = 0x9F; else if (c == 0xF0) lo = 0x90; else if (c == 0xF4) hi = 0x8F;
    if (p[1] < lo || p[1] > hi) return 0;
    for (size_t k = 2; k <= need; ++k) if ((p[k] & 0xC0) != 0x80) return 0;
    return need + 1;
  }

  bool hex4(uint32_t& v) {
    if (e_ - p_ < 4) return fail("short \\u escape");
    v = 0;
    for (int k = 0; k < 4; ++k) {
      char c = *p_++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
      else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
      else return fail("bad \\u escape");
    }
    return true;
  }

  bool string(std::string& out) {
    if (p_ >= e_ || *p_ != '"') return fail("expected string");
    ++p_;
    out.clear();
    while (true) {
      if (p_ >= e_) return fail("unterminated string");
      unsigned char c = (unsigned char)*p_++;
      if (c == '"') return true;
      if (c < 0x20) return fail("control character in string");
      if (c != '\\') {
        if (c < 0x80) { out.push_back(char(c)); continue; }
        // encoding/json coerces string contents to well-formed UTF-8: every byte that does not start a valid sequence
        // becomes U+FFFD (utf8.DecodeRune semantics, one replacement per byte)
        --p_;
        size_t w = utf8_len(p_, size_t(e_ - p_));
        if (w == 0) { put_utf8(out, 0xFFFD); ++p_; }
        else { out.append(p_, w); p_ += w; }
        continue;
      }
      if (p_ >= e_) return fail("unterminated escape");
      char esc = *p_++;
      switch (esc) {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
          uint32_t u;
          if (!hex4(u)) return false;
          if (u >= 0xD800 && u <= 0xDBFF) {   // high surrogate: need a low one
            if (e_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
              const char* save = p_;
              p_ += 2;
              uint32_t lo;
              if (!hex4(lo)) return false;
              if (lo >= 0xDC00 && lo <= 0xDFFF) u = 0x10000 + ((u - 0xD800) << 10) + (lo - 0xDC00);
              else { p_ = save; u = 0xFFFD; }
            } else u = 0xFFFD;
          } else if (u >= 0xDC00 && u <= 0xDFFF) u = 0xFFFD;
          put_utf8(out, u);
          break;
        }
        default: return fail("bad escape");
      }
    }
  }

  bool number(Value& v) {
    const char* s = p_;
    if (p_ < e_ && *p_ == '-') ++p_;
    if (p_ >= e_) return fail("bad number");
    if (*p_ == '0') ++p_;
    else if (*p_ >= '1' && *p_ <= '9') { while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_; }
    else return fail("bad number");
    bool integral = true;
    if (p_ < e_ && *p_ == '.') {
      integral = false; ++p_;
      if (p_ >= e_ || *p_ < '0' || *p_ > '9') return fail("bad fraction");
      while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
    }
    if (p_ < e_ && (*p_ == 'e' || *p_ == 'E')) {
      integral = false; ++p_;
      if (p_ < e_ && (*p_ == '+' || *p_ == '-')) ++p_;
      if (p_ >= e_ || *p_ < '0' || *p_ > '9') return fail("bad exponent");
      while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
    }
    std::string tmp(s, size_t(p_ - s));
    v.kind = Kind::Number;
    v.d = std::strtod(tmp.c_str(), nullptr);
    v.is_int = false;
    if (integral && tmp.size() <= 19) {
      char* end = nullptr;
      errno = 0;
      long long ll = std::strtoll(tmp.c_str(), &end, 10);
      if (end && *end == 0 && errno != ERANGE) { v.is_int = true; v.i = ll; }   // 2^63 saturates: keep the double
    }
    return true;
  }

  bool value(Value& v, int depth) {
    if (depth > 256) return fail("nesting too deep");
    if (p_ >= e_) return fail("unexpected end of input");
    char c = *p_;
    if (c == '{') {
      ++p_;
      v.kind = Kind::Object;
      ws();
      if (p_ < e_ && *p_ == '}') { ++p_; return true; }
      while (true) {
        ws();
        std::string k;
        if (!string(k)) return false;
        ws();
        if (p_ >= e_ || *p_ != ':') return fail("expected ':'");
        ++p_;
        ws();
        v.obj.emplace_back(std::move(k), Value{});
        if (!value(v.obj.back().second, depth + 1)) return false;
        ws();
        if (p_ < e_ && *p_ == ',') { ++p_; continue; }
        if (p_ < e_ && *p_ == '}') { ++p_; return true; }
        return fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      ++p_;
      v.kind = Kind::Array;
      ws();
      if (p_ < e_ && *p_ == ']') { ++p_; return true; }
      while (true) {
        ws();
        v.arr.emplace_back();
        if (!value(v.arr.back(), depth + 1)) return false;
        ws();
        if (p_ < e_ && *p_ == ',') { ++p_; continue; }
        if (p_ < e_ && *p_ == ']') { ++p_; return true; }
        return fail("expected ',' or ']'");
      }
    }
    if (c == '"') { v.kind = Kind::String; return string(v.s); }
    if (c == 't') { v.kind = Kind::Bool; v.b = true; return lit("true"); }
    if (c == 'f') { v.kind = Kind::Bool; v.b = false; return lit("false"); }
    if (c == 'n') { v.kind = Kind::Null; return lit("null"); }
    if (c == '-' || (c >= '0' && c <= '9')) return number(v);
    return fail("unexpected character");
  }
};

inline bool parse(std::string_view text, Value& out, std::string* err = nullptr) {
  Parser p(text.data(), text.size());
  return p.parse(out, err);
}

}  // namespace mjson
