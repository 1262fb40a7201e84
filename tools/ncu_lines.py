"""Per-source-line instruction / stall-sample breakdown of one kernel in an ncu report.
usage: python tools/ncu_lines.py report.ncu-rep '<demangled-name-substring>' '<mangled-substring>' [top]"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, dem, mang = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 24
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = None
    for k, i in enumerate(starts):
        if dem in rows[i][1]:
            sec = (i, starts[k + 1] if k + 1 < len(starts) else len(rows))
            break
    if sec is None:
        sys.exit("kernel not found")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "cordum_b200", "libcordum_b200.so")], cwd=d, capture_output=True)
        cub = glob.glob(os.path.join(d, "kernels*.cubin"))[0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    amap, line, fn = {}, None, None
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and fn and mang in fn:
            amap[int(m.group(1), 16)] = line
    h = rows[sec[0] + 1]
    ai, ns, ie, te, si = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Source")
    src = open(os.path.join(ROOT, "cordum_b200", "csrc", "kernels.cu")).read().splitlines()
    base, agg, ops, tot, tots = None, {}, {}, 0, 0
    for r in rows[sec[0] + 2: sec[1]]:
        try:
            a, n, s, t = int(r[ai], 16), int(r[ie] or 0), int(r[ns] or 0), int(r[te] or 0)
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        ln = amap.get(a - base)
        d = agg.setdefault(ln, [0, 0, 0])
        d[0] += n
        d[1] += s
        d[2] += t
        toks = r[si].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + n
        tot += n
        tots += s
    print("total warp-instructions", tot, "samples", tots)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[k[1] - 1].strip()[:96] if k and k[0] == "kernels.cu" else str(k)
        print("%5.1f%% inst %5.1f%% samp thr %4.1f  L%s %s" % (100 * v[0] / tot, 100 * v[1] / max(tots, 1), v[2] / max(v[0], 1), k[1] if k else "?", text))
    print("opcodes:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]))


if __name__ == "__main__":
    main()
