"""Condenses one ncu launch list (gpurun_out/launches_<name>.csv, written by tools/gpu_launchlist.sh) into
profiles/<tag>_<name>_launch_summary.md.

    python tools/summarize_launches.py r01j pan "one 800x1000 frame (63 tiles), 4x PAN nf=40 unf=24 nb=16 fp16"
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name, what = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % name)) if not l.startswith("==")]
agg, seq, total = collections.OrderedDict(), [], 0.0
for r in csv.DictReader(lines):
    k = r["Kernel Name"].split("(")[0].replace("void innfer::<unnamed>::", "").replace("innfer::<unnamed>::", "")
    us = float(r["Metric Value"].replace(",", "")) / {"ns": 1e3, "us": 1.0, "ms": 1e-3}[r["Metric Unit"]]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    seq.append((k, us))
out = os.path.join(ROOT, "profiles", "%s_%s_launch_summary.md" % (tag, name))
with open(out, "w") as f:
    f.write("# ncu launch list (%s): %s\n\n" % (tag, what))
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over every launch "
            "(cold-cache, serialised: compare shares).\n\n")
    f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f | %.2f%% |\n" % (k, n, us, us / n, 100 * us / total))
    f.write("\ntotal device time %.1f ms over %d launches\n\n" % (total / 1e3, len(seq)))
    f.write("Launch order (us): " + " ".join("%.0f" % us for _, us in seq) + "\n")
print(open(out).read()[:1500])
