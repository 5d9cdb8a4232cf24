"""Condenses gpurun_out/launches.csv (ncu gpu__time_duration pass) and gpurun_out/prof_conv.ncu-rep
(ncu --set full on a few conv launches) into small tracked files under profiles/.

    python tools/summarize_profile.py r01b
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[unit]


# ---- launch list
lines = [l for l in open(os.path.join(ROOT, "gpurun_out", "launches.csv")) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
total = 0.0
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void innfer::<unnamed>::", "")
    us = to_us(r["Metric Value"], r["Metric Unit"])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
with open(os.path.join(out_dir, "%s_launch_summary.md" % tag), "w") as f:
    f.write("# ncu launch list (%s): one 800x1000 frame (63 tiles, one batch), 4x RRDB nb=23 fp16\n\n" % tag)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over every launch of "
            "`tests/gpu_bringup.py --stage prof` (cold-cache, serialised: compare shares).\n\n")
    f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f | %.2f%% |\n" % (name, n, t, t / n, 100 * t / total))
    f.write("\ntotal device time %.1f ms over %d launches\n\n" % (total / 1e3, len(rows)))
    f.write("First RRDB of the first batch, in launch order (conv1..conv5 of RDB1, ...):\n\n| # | kernel | us |\n|---|---|---:|\n")
    for i, r in enumerate(rows[1:17]):
        f.write("| %d | `%s` | %.1f |\n" % (i, r["Kernel Name"].split("(")[0][-22:], to_us(r["Metric Value"], r["Metric Unit"])))

# ---- full capture of a few conv launches
rep = os.path.join(ROOT, "gpurun_out", "prof_conv.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
want = [w for w in want if w in idx]
with open(os.path.join(out_dir, "%s_conv_ncu_full.csv" % tag), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(data))])
    for m in want:
        w.writerow([m, units[idx[m]]] + [d[idx[m]][:60] for d in data])


def fbytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


tr = []
for d in data:
    rd = fbytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = fbytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    tr.append(rd + wr)
with open(os.path.join(out_dir, "conv_traffic.json"), "w") as f:
    json.dump({"source": "%s_conv_ncu_full.csv (ncu --set full, %d consecutive conv launches of one RDB cycle, "
                         "one batch of 63 tiles = 800x1000 frame)" % (tag, len(data)),
               "dram_bytes_per_launch": tr, "dram_bytes_per_launch_avg": sum(tr) / len(tr)}, f, indent=1)
print(open(os.path.join(out_dir, "%s_launch_summary.md" % tag)).read()[:1500])
print(open(os.path.join(out_dir, "%s_conv_ncu_full.csv" % tag)).read()[:3000])
