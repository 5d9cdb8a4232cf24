"""Launch sequence of an ncu csv log with time, DRAM bytes and tensor-pipe activity per launch.
usage: python tools/ncu_seq.py LOG.csv [first [count]]"""
import collections
import csv
import sys


def load(f):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = collections.OrderedDict()
    for r in rows[1:]:
        k = r[iK].split("(")[0].replace("void innfer::<unnamed>::", "").replace("innfer::<unnamed>::", "")
        d.setdefault((int(r[iID]), k), {})[r[iM]] = float(r[iV].replace(",", ""))
    return d


def main():
    d = load(sys.argv[1])
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    count = int(sys.argv[3]) if len(sys.argv) > 3 else len(d)
    print("total %.2f ms over %d launches" % (sum(m["gpu__time_duration.sum"] for m in d.values()) / 1e6, len(d)))
    agg = collections.OrderedDict()
    for (i, k), m in d.items():
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += m["gpu__time_duration.sum"] / 1e3
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-48s n=%4d total %9.1f us avg %8.1f" % (k, n, us, us / n))
    for (i, k), m in list(d.items())[first:first + count]:
        t = m["gpu__time_duration.sum"] / 1e3
        rd, wr = m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6
        pipe = m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
        print("%4d %-45s %7.1f us rd %7.1f wr %7.1f MB %5.0f GB/s pipe %4.1f%%" % (i, k, t, rd, wr, (rd + wr) / t * 1e3, pipe))


if __name__ == "__main__":
    main()
