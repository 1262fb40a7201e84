"""Summaries committed under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).
  python tools/profile_summary.py launches <ncu --csv launch list> <out.json>
  python tools/profile_summary.py full <report.ncu-rep> <out.json>
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[k], [0, 0.0])
        a[0] += 1
        a[1] += float(r[v].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    res = {name: {"launches": a[0], "mean_us": round(a[1] / a[0] / 1e3, 2), "share_of_gpu_time": round(a[1] / total, 4)}
           for name, a in agg.items()}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = OrderedDict()
        d["Kernel Name"] = r[hdr.index("Kernel Name")]
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                d[m] = ("%s %s" % (r[i], units[i])).strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        d["top_stalls_warps_per_issue_active"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
