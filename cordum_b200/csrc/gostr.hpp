// gostr.hpp — the string semantics the table compiler needs, written for the PRODUCT.
//
// All string work of the path happens once per dictionary value at table-compile /
// encode time, never on the device.  These helpers restate the Go
// stdlib behaviour the reference relies on:
//   strings.TrimSpace   (kernel.go:133-134, safety_policy.go:301,357)
//   strings.EqualFold   (safety_policy.go:301)   -> canonical form = every rune replaced by the representative of
//                       its simple-case-folding class (all of Unicode 15.0.0, as go1.24: U+212A KELVIN ~ k,
//                       U+017F ~ s, final sigma ~ sigma; U+0130 / U+0131 fold with nothing), invalid UTF-8 -> U+FFFD
//   strings.ToLower     (strategy_least_loaded.go:250,256; kernel.go:403) -> unicode.ToLower per rune (a different
//                       canonical form: U+017F lower-cases to itself)
//   path.Match          (safety_policy.go:361; kernel.go:482)
//   fmt %q / strconv.Quote  (safety_policy.go:410,413 MCP reasons, :235 tenant reasons; gateway/policy_bundles.go:1207,1211)
// Implementation is independent of oracle/ (different algorithms on purpose): globs are
// compiled once into element lists and matched with single-star backtracking.
// Tables: common/go_unicode_tables.h (generated from the Unicode Character Database, tools/gen_go_unicode.py).
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>

#include "../../common/go_unicode_tables.h"

namespace cordum {

using sv = std::string_view;

// ---- UTF-8 (Go's utf8.DecodeRune semantics: invalid byte -> U+FFFD, width 1)
inline uint32_t utf8_next(const unsigned char* p, size_t n, int& w) {
  if (n == 0) { w = 0; return 0xFFFD; }
  unsigned c = p[0];
  w = 1;
  if (c < 0x80) return c;
  if (c < 0xC2 || c > 0xF4) return 0xFFFD;
  size_t need = (c >= 0xF0) ? 3 : (c >= 0xE0) ? 2 : 1;   // continuation bytes
  if (n < need + 1) return 0xFFFD;
  unsigned lo = 0x80, hi = 0xBF;
  if (c == 0xE0) lo = 0xA0; else if (c == 0xED) hi = 0x9F; else if (c == 0xF0) lo = 0x90; else if (c == 0xF4) hi = 0x8F;
  if (p[1] < lo || p[1] > hi) return 0xFFFD;
  uint32_t r = (need == 1) ? (c & 0x1F) : (need == 2) ? (c & 0x0F) : (c & 0x07);
  r = (r << 6) | (p[1] & 0x3F);
  for (size_t k = 2; k <= need; ++k) {
    if ((p[k] & 0xC0) != 0x80) return 0xFFFD;
    r = (r << 6) | (p[k] & 0x3F);
  }
  w = (int)need + 1;
  return r;
}

inline bool go_is_space(uint32_t r) {
  switch (r) {
    case 0x09: case 0x0A: case 0x0B: case 0x0C: case 0x0D: case 0x20: case 0x85: case 0xA0:
    case 0x1680: case 0x2028: case 0x2029: case 0x202F: case 0x205F: case 0x3000:
      return true;
    default:
      return r >= 0x2000 && r <= 0x200A;
  }
}

// strings.TrimSpace.  The right side mirrors utf8.DecodeLastRuneInString on the left-trimmed
// string: look back over at most 3 bytes for a rune start; the rune counts only if it is whole
// and ends exactly at the tail.
inline sv trim_space(sv s) {
  const unsigned char* p = (const unsigned char*)s.data();
  size_t a = 0, b = s.size();
  while (a < b) {
    int w;
    uint32_t r = utf8_next(p + a, b - a, w);
    if (!go_is_space(r)) break;
    a += (size_t)w;
  }
  while (b > a) {
    unsigned c = p[b - 1];
    if (c < 0x80) {
      if (!go_is_space(c)) break;
      --b;
      continue;
    }
    size_t lim = (b - a >= 4) ? b - 4 : a;
    size_t start = b;   // sentinel: none found
    for (size_t k = b - 1; k > lim;) {
      --k;
      if ((p[k] & 0xC0) != 0x80) { start = k; break; }
    }
    if (start == b) break;
    int w;
    uint32_t r = utf8_next(p + start, b - start, w);
    if (start + (size_t)w != b || !go_is_space(r)) break;
    b = start;
  }
  return s.substr(a, b - a);
}

inline char lower_ascii(char c) { return (c >= 'A' && c <= 'Z') ? char(c + 32) : c; }

inline uint32_t rune_map(const GoRunePair* tab, uint32_t n, uint32_t r) {   // sorted by .from
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (tab[mid].from < r) lo = mid + 1; else hi = mid;
  }
  return (lo < n && tab[lo].from == r) ? tab[lo].to : r;
}
inline void utf8_append(std::string& o, uint32_t r) {
  if (r < 0x80) o.push_back((char)r);
  else if (r < 0x800) { o.push_back((char)(0xC0 | (r >> 6))); o.push_back((char)(0x80 | (r & 0x3F))); }
  else if (r < 0x10000) { o.push_back((char)(0xE0 | (r >> 12))); o.push_back((char)(0x80 | ((r >> 6) & 0x3F))); o.push_back((char)(0x80 | (r & 0x3F))); }
  else { o.push_back((char)(0xF0 | (r >> 18))); o.push_back((char)(0x80 | ((r >> 12) & 0x3F))); o.push_back((char)(0x80 | ((r >> 6) & 0x3F))); o.push_back((char)(0x80 | (r & 0x3F))); }
}
inline bool is_ascii(sv s) {
  unsigned char acc = 0;
  for (unsigned char c : s) acc |= c;
  return acc < 0x80;
}
// rune-wise map of a string (Go: `for _, r := range s`: an invalid byte decodes to U+FFFD, width 1)
inline std::string map_runes(sv s, const GoRunePair* tab, uint32_t n) {
  std::string o;
  o.reserve(s.size());
  const unsigned char* p = (const unsigned char*)s.data();
  for (size_t i = 0; i < s.size();) {
    int w;
    uint32_t r = utf8_next(p + i, s.size() - i, w);
    i += (size_t)w;
    utf8_append(o, r < 0x80 ? (uint32_t)(unsigned char)lower_ascii((char)r) : rune_map(tab, n, r));
  }
  return o;
}

// Canonical form under strings.EqualFold(a, b): two strings are EqualFold iff their fold_str are byte-equal.
inline std::string fold_str(sv s) {
  if (is_ascii(s)) { std::string o(s); for (auto& c : o) c = lower_ascii(c); return o; }
  return map_runes(s, kGoFoldRep, kGoFoldRepCount);
}
// Canonical form under strings.EqualFold(TrimSpace(a), TrimSpace(b)).
inline std::string fold_key(sv s) { return fold_str(trim_space(s)); }
// strings.ToLower
inline std::string lower_copy(sv s) {
  if (is_ascii(s)) { std::string o(s); for (auto& c : o) c = lower_ascii(c); return o; }
  return map_runes(s, kGoLower, kGoLowerCount);
}
// strings.ToLower(strings.TrimSpace(s))
inline std::string lower_key(sv s) { return lower_copy(trim_space(s)); }

// strconv.IsPrint: letters, marks, numbers, punctuation, symbols and the ASCII space (Unicode 15.0.0 as go1.24).
inline bool go_is_print(uint32_t r) {
  if (r < 0x7F) return r >= 0x20;
  uint32_t lo = 0, hi = kGoPrintCount;   // first range whose end is >= r
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (kGoPrint[mid].to < r) lo = mid + 1; else hi = mid;
  }
  return lo < kGoPrintCount && kGoPrint[lo].from <= r;
}
// strconv.Quote (what fmt's %q prints for a string): printable runes pass, `"` and `\` get a backslash, the seven C
// escapes are named, other control bytes and undecodable bytes are \xhh, the remaining runes \uhhhh / \Uhhhhhhhh.
inline std::string go_quote(sv s) {
  static const char* hex = "0123456789abcdef";
  std::string o = "\"";
  const unsigned char* p = (const unsigned char*)s.data();
  for (size_t i = 0; i < s.size();) {
    int w;
    uint32_t r = utf8_next(p + i, s.size() - i, w);
    if (w == 1 && r == 0xFFFD) {   // a byte that starts no valid sequence
      o += "\\x"; o.push_back(hex[p[i] >> 4]); o.push_back(hex[p[i] & 15]);
      ++i;
      continue;
    }
    i += (size_t)w;
    if (r == '"' || r == '\\') { o.push_back('\\'); o.push_back((char)r); continue; }
    if (go_is_print(r)) { utf8_append(o, r); continue; }
    switch (r) {
      case 7: o += "\\a"; continue;
      case 8: o += "\\b"; continue;
      case 12: o += "\\f"; continue;
      case 10: o += "\\n"; continue;
      case 13: o += "\\r"; continue;
      case 9: o += "\\t"; continue;
      case 11: o += "\\v"; continue;
      default: break;
    }
    if (r < 0x20 || r == 0x7F) { o += "\\x"; o.push_back(hex[r >> 4]); o.push_back(hex[r & 15]); }
    else if (r < 0x10000) { o += "\\u"; for (int sh = 12; sh >= 0; sh -= 4) o.push_back(hex[(r >> sh) & 15]); }
    else { o += "\\U"; for (int sh = 28; sh >= 0; sh -= 4) o.push_back(hex[(r >> sh) & 15]); }
  }
  o.push_back('"');
  return o;
}
inline bool starts_with(sv s, sv p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }

// ------------------------------------------------------------------ compiled glob
// path.Match semantics:
//   '*'  any run of non-'/' bytes            '?'  one rune, not '/'
//   [..] one rune in (or, with ^, not in) the ranges; '\\' escapes; malformed -> never matches
// A malformed pattern is malformed for every name (Go validates the rest of the pattern even
// after a failed chunk), so validity is a property of the compiled object.
class Glob {
 public:
  explicit Glob(sv pattern) { valid_ = compile(pattern); }
  bool valid() const { return valid_; }

  bool match(sv name) const {
    if (!valid_) return false;
    const unsigned char* s = (const unsigned char*)name.data();
    const size_t n = name.size();
    size_t si = 0, ei = 0;
    size_t star_e = (size_t)-1, star_s = 0;   // last star: element after it, and how far it has eaten
    while (true) {
      if (ei < elems_.size() && elems_[ei].kind == kStar) {
        star_e = ++ei;
        star_s = si;
        continue;
      }
      if (ei == elems_.size()) {
        if (si == n) return true;
      } else if (si < n) {
        int w = step(elems_[ei], s + si, n - si);
        if (w > 0) { si += (size_t)w; ++ei; continue; }
      }
      // mismatch: let the last star eat one more byte (never a '/')
      if (star_e == (size_t)-1 || star_s >= n || s[star_s] == '/') return false;
      ++star_s;
      si = star_s;
      ei = star_e;
    }
  }

 private:
  enum Kind : uint8_t { kLit, kAny, kClass, kStar };
  struct Range { uint32_t lo, hi; };
  struct Elem { Kind kind; uint8_t lit; bool neg; uint32_t r0, r1; };   // class: ranges_[r0, r1)
  std::vector<Elem> elems_;
  std::vector<Range> ranges_;
  bool valid_ = false;

  // width consumed if the element accepts the text at p, else 0
  int step(const Elem& e, const unsigned char* p, size_t n) const {
    if (e.kind == kLit) return p[0] == e.lit ? 1 : 0;
    int w;
    uint32_t r = utf8_next(p, n, w);
    if (e.kind == kAny) return p[0] == '/' ? 0 : w;
    bool in = false;
    for (uint32_t k = e.r0; k < e.r1; ++k)
      if (ranges_[k].lo <= r && r <= ranges_[k].hi) { in = true; break; }
    return in != e.neg ? w : 0;
  }

  // one class endpoint (getEsc): false = malformed
  static bool endpoint(const unsigned char* p, size_t n, size_t& i, uint32_t& out) {
    if (i >= n || p[i] == '-' || p[i] == ']') return false;
    if (p[i] == '\\') { if (++i >= n) return false; }
    int w;
    out = utf8_next(p + i, n - i, w);
    if (out == 0xFFFD && w == 1) return false;
    i += (size_t)w;
    return i < n;   // a class can never end the pattern without its ']'
  }

  bool compile(sv pat) {
    const unsigned char* p = (const unsigned char*)pat.data();
    const size_t n = pat.size();
    size_t i = 0;
    while (i < n) {
      unsigned char c = p[i];
      if (c == '*') {
        if (elems_.empty() || elems_.back().kind != kStar) elems_.push_back({kStar, 0, false, 0, 0});
        ++i;
      } else if (c == '?') {
        elems_.push_back({kAny, 0, false, 0, 0});
        ++i;
      } else if (c == '[') {
        ++i;
        Elem e{kClass, 0, false, (uint32_t)ranges_.size(), 0};
        if (i < n && p[i] == '^') { e.neg = true; ++i; }
        int nrange = 0;
        while (true) {
          if (i < n && p[i] == ']' && nrange > 0) { ++i; break; }
          uint32_t lo, hi;
          if (!endpoint(p, n, i, lo)) return false;
          hi = lo;
          if (p[i] == '-') {
            ++i;
            if (!endpoint(p, n, i, hi)) return false;
          }
          ranges_.push_back({lo, hi});
          ++nrange;
        }
        e.r1 = (uint32_t)ranges_.size();
        elems_.push_back(e);
      } else {
        if (c == '\\') {
          if (++i >= n) return false;
          c = p[i];
        }
        elems_.push_back({kLit, c, false, 0, 0});
        ++i;
      }
    }
    return true;
  }
};

}  // namespace cordum
