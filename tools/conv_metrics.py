"""ncu CSV (tools/r02_housekeeping.sh) -> profiles/r02_conv_metrics.json: per conv kernel family of the RRDB trunk the
per-launch averages of DRAM bytes, duration, L2 hit rate and tensor-pipe activity, measured at the bench's own batch size."""
import collections
import csv
import json
import sys

# template arguments of conv_rows_kernel<COUT, KSLABS, RES, PAIR, ...> -> family names used by bench.py
FAMILIES = {"<32, 4, 0, 0": "conv_rows 64->32", "<32, 6, 0, 0": "conv_rows 96->32", "<32, 8, 0, 0": "conv_rows 128->32",
            "<32, 5, 0, 0": "conv_rows 160->32", "<64, 6, 1, 1": "conv_rows_pair 192->64 +res",
            # round 2e: conv5 of RDB1 / RDB2 takes its block residual through identity MMAs and runs the RES = 0 pair
            # instantiation; bench.py keeps one family for all three conv5s of an RRDB (the capture must hold whole
            # RRDBs -- a multiple of 15 launches -- for the average to weigh them 2 : 1 as a step does)
            "<64, 6, 0, 1": "conv_rows_pair 192->64 +res"}

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
by_id = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = by_id.setdefault(row["ID"], {"name": row["Kernel Name"], "grid": row["Grid Size"]})
    d[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
fam = collections.OrderedDict()
for d in by_id.values():
    name = None
    for key, val in FAMILIES.items():
        if "conv_rows_kernel" + key in d["name"].replace("(int)", ""):
            name = val
    if name is None:
        name = d["name"].split("::")[-1].split("(")[0]
    fam.setdefault(name, []).append(d)
out = {"source": "ncu --cache-control none --clock-control none over `bench.py --steps 1 --warmup 3` (1080p frame, batches of 95 "
                 "tiles), 45 consecutive conv launches of the trunk (three RRDBs); per-launch averages", "families": {}}
for name, ds in fam.items():
    n = len(ds)
    avg = lambda k: sum(d.get(k, 0.0) for d in ds) / n
    out["families"][name] = {
        "launches_captured": n,
        "dram_bytes_per_launch": avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum"),
        "dram_read_bytes": avg("dram__bytes_read.sum"), "dram_write_bytes": avg("dram__bytes_write.sum"),
        "duration_us": avg("gpu__time_duration.sum") / 1e3,
        "l2_hit_pct": avg("lts__t_sector_hit_rate.pct"),
        "tensor_pipe_active_pct": avg("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    }
dom = max(out["families"].items(), key=lambda kv: kv[1]["duration_us"] * kv[1]["launches_captured"])
out["dominant"] = dict(dom[1], kernel=dom[0])
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, v in out["families"].items():
    print("%-32s n=%2d %8.1f us  dram %8.1f MB  L2 hit %5.1f%%  pipe %5.1f%%" %
          (k, v["launches_captured"], v["duration_us"], v["dram_bytes_per_launch"] / 1e6, v["l2_hit_pct"], v["tensor_pipe_active_pct"]))
