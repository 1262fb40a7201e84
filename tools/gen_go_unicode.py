#!/usr/bin/env python
"""Generates common/go_unicode_tables.h: the Unicode 15.0.0 data behind Go 1.24's strings.EqualFold / strings.ToLower.

Source of the data: the Unicode Character Database shipped with this image's perl (Unicode::UCD, unicore version
15.0.0 == go1.24's unicode.Version), NOT the reference repository.  What Go does with it (restated from the go1.24
documentation of unicode.SimpleFold and strings.EqualFold):

  SimpleFold(r): the smallest rune > r that is equivalent to r under Unicode simple case folding, or, if there is
                 none, the smallest rune >= 0 equivalent to r (orbits are walked in ascending cyclic order);
                 runes outside any multi-element folding orbit fall back to ToLower(r), then ToUpper(r).
  EqualFold:     rune by rune; for a differing pair, with sr < tr: ASCII tr -> only 'A'-'Z' against 'a'-'z';
                 otherwise walk r = SimpleFold(sr) while r != sr and r < tr, equal iff r == tr.
  ToLower:       per rune, the simple lowercase mapping (UnicodeData.txt field 13); invalid UTF-8 -> U+FFFD.

The generator emulates that walk for every pair of related runes and emits
  kGoFoldRep[]   rune -> canonical representative of its EqualFold class (only where rep != rune)
  kGoLower[]     rune -> simple lowercase mapping (only where != rune)
  kGoPrint[]     rune ranges strconv.Quote leaves unescaped (strconv.IsPrint)
It also asserts that the emulated EqualFold relation is an equivalence on every class it emits (the one-directional
mappings of U+0130 / U+0131 must relate nothing), so "same representative" is exactly "EqualFold".
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PERL = r'''
use Unicode::UCD qw(prop_invmap);
use strict;
print "{";
my $first = 1;
for my $p ("Simple_Case_Folding", "Simple_Lowercase_Mapping", "Simple_Uppercase_Mapping") {
  my ($list, $map, $fmt, $def) = prop_invmap($p);
  die "unexpected format $fmt" unless $fmt eq "a" && $def eq "0";
  print "," unless $first; $first = 0;
  print "\"$p\":[";
  my $f2 = 1;
  for (my $i = 0; $i < @$list; $i++) {
    my $lo = $list->[$i];
    my $hi = ($i + 1 < @$list) ? $list->[$i + 1] - 1 : 0x10FFFF;
    my $m = $map->[$i];
    next if $m == 0;
    print "," unless $f2; $f2 = 0;
    print "[$lo,$hi,$m]";
  }
  print "]";
}
{
  my ($list, $map, $fmt, $def) = prop_invmap("General_Category");
  print ",\"gc\":[";
  for (my $i = 0; $i < @$list; $i++) {
    my $lo = $list->[$i];
    my $hi = ($i + 1 < @$list) ? $list->[$i + 1] - 1 : 0x10FFFF;
    print "," if $i;
    print "[$lo,$hi,\"" . $map->[$i] . "\"]";
  }
  print "]";
}
print ",\"version\":\"" . Unicode::UCD::UnicodeVersion() . "\"}";
'''


def load():
    out = subprocess.run(["perl", "-e", PERL], check=True, capture_output=True, text=True).stdout
    d = json.loads(out)
    assert d["version"] == "15.0.0", d["version"]

    def expand(name):
        m = {}
        for lo, hi, base in d[name]:
            for cp in range(lo, hi + 1):
                m[cp] = base + (cp - lo)
        return m
    global GC
    GC = d["gc"]
    return expand("Simple_Case_Folding"), expand("Simple_Lowercase_Mapping"), expand("Simple_Uppercase_Mapping")


GC = None


def print_ranges():
    """unicode.IsPrint / strconv.IsPrint: categories L, M, N, P, S plus U+0020 (go1.24 unicode/graphic.go: "letters,
    marks, numbers, punctuation, symbols, and the ASCII space character").  Merged inclusive ranges."""
    out = []
    for lo, hi, cat in GC:
        ok = cat[0] in "LMNPS"
        if lo <= 0x20 <= hi and not ok:      # the ASCII space sits in a Zs range of its own
            assert (lo, hi) == (0x20, 0x20), (lo, hi, cat)
            ok = True
        if not ok:
            continue
        if out and out[-1][1] + 1 == lo:
            out[-1][1] = hi
        else:
            out.append([lo, hi])
    return out


def main():
    scf, slc, suc = load()
    lower = lambda r: slc.get(r, r)
    upper = lambda r: suc.get(r, r)
    # folding orbits: runes with the same simple case folding
    classes = {}
    for r in set(scf) | set(scf.values()):
        classes.setdefault(scf.get(r, r), set()).add(r)
    orbit_next = {}
    for members in classes.values():
        ms = sorted(members)
        if len(ms) < 2:
            continue
        # Go keeps an orbit table entry only when the class is not the plain {ToLower, ToUpper} pair; the walk below
        # is the same either way because for a plain pair lower/upper give exactly the cyclic successor.
        for i, r in enumerate(ms):
            orbit_next[r] = ms[(i + 1) % len(ms)]
    plain_pairs = 0

    def simple_fold(r):
        if r in orbit_next:
            return orbit_next[r]
        lo = lower(r)
        if lo != r:
            return lo
        return upper(r)

    # check: inside a folding orbit, lower()/upper() fallbacks agree with the cyclic successor for 2-element orbits
    for members in classes.values():
        ms = sorted(members)
        if len(ms) == 2:
            a, b = ms
            la = lower(a) if lower(a) != a else upper(a)
            lb = lower(b) if lower(b) != b else upper(b)
            if la == b and lb == a:
                plain_pairs += 1

    def equal_fold_rune(sr, tr):
        if sr == tr:
            return True
        if tr < sr:
            sr, tr = tr, sr
        if tr < 0x80:
            return 0x41 <= sr <= 0x5A and tr == sr + 0x20
        r = simple_fold(sr)
        while r != sr and r < tr:
            r = simple_fold(r)
        return r == tr

    # every rune that any mapping touches
    touched = set(scf) | set(scf.values()) | set(slc) | set(slc.values()) | set(suc) | set(suc.values())
    # relation graph over touched runes: candidates related to r are in its orbit or reachable by repeated SimpleFold
    rep = {}
    related = {}
    for r in sorted(touched):
        seen, x = [r], simple_fold(r)
        for _ in range(8):
            if x in seen:
                break
            seen.append(x)
            x = simple_fold(x)
        related[r] = seen
    groups = {}
    for r in sorted(touched):
        eq = sorted({x for x in related[r] if equal_fold_rune(r, x)} | {r})
        groups[r] = eq
    # equivalence check: symmetric + transitive on what we emit
    for r, eq in groups.items():
        for x in eq:
            assert r in groups.get(x, [x]) or x == r, "asymmetric fold relation U+%04X U+%04X" % (r, x)
            assert groups.get(x, [x]) == eq or x == r and len(eq) == 1, "not an equivalence at U+%04X" % r
    # U+0130 / U+0131 must be related to nothing (their lower/upper mappings are one-directional)
    assert groups[0x130] == [0x130] and groups[0x131] == [0x131]
    assert equal_fold_rune(0x212A, ord("k")) and equal_fold_rune(0x212A, ord("K")) and equal_fold_rune(0x17F, ord("S"))
    assert equal_fold_rune(0x3C2, 0x3A3) and equal_fold_rune(0x3C2, 0x3C3)
    for r, eq in groups.items():
        if len(eq) < 2:
            continue
        ascii_lower = [x for x in eq if 0x61 <= x <= 0x7A]
        # representative: the ASCII lower-case letter if the class has one (keeps the ASCII canonical form), else the
        # simple lowercase mapping of the smallest member if it is in the class, else the smallest member
        if ascii_lower:
            c = ascii_lower[0]
        else:
            lo = lower(eq[0])
            c = lo if lo in eq else eq[0]
        if c != r:
            rep[r] = c
    lowers = {r: v for r, v in slc.items() if v != r}
    path = os.path.join(ROOT, "common", "go_unicode_tables.h")
    with open(path, "w") as f:
        f.write("// go_unicode_tables.h - GENERATED by tools/gen_go_unicode.py from the Unicode Character Database 15.0.0\n"
                "// (perl Unicode::UCD of this image; go1.24's unicode.Version is 15.0.0).  Do not edit.\n"
                "//   kGoFoldRep: rune -> representative of its strings.EqualFold class (entries only where rep != rune)\n"
                "//   kGoLower:   rune -> unicode.ToLower (simple lowercase mapping; entries only where != rune)\n"
                "#pragma once\n#include <stdint.h>\n\n"
                "struct GoRunePair { uint32_t from, to; };\n\n")
        for name, table in (("kGoFoldRep", rep), ("kGoLower", lowers)):
            items = sorted(table.items())
            f.write("static const GoRunePair %s[%d] = {\n" % (name, len(items)))
            for i in range(0, len(items), 6):
                f.write("  " + " ".join("{0x%X,0x%X}," % kv for kv in items[i:i + 6]) + "\n")
            f.write("};\nstatic const uint32_t %sCount = %d;\n\n" % (name, len(items)))
        pr = print_ranges()
        f.write("// kGoPrint: inclusive rune ranges for which strconv.IsPrint is true (categories L, M, N, P, S and U+0020)\n"
                "static const GoRunePair kGoPrint[%d] = {\n" % len(pr))
        for i in range(0, len(pr), 6):
            f.write("  " + " ".join("{0x%X,0x%X}," % (a, b) for a, b in pr[i:i + 6]) + "\n")
        f.write("};\nstatic const uint32_t kGoPrintCount = %d;\n" % len(pr))
    print("wrote %s: %d fold entries, %d lower entries (%d plain 2-orbits)" % (path, len(rep), len(lowers), plain_pairs))


if __name__ == "__main__":
    main()
