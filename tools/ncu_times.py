"""Print per-launch time / DRAM bytes from an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum` log.  usage: python tools/ncu_times.py LOG.csv [substring ...]"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((r[iID], r[iK][:70]), {})[r[iM]] = float(r[iV].replace(",", ""))
    pats = sys.argv[2:]
    for (i, k), m in d.items():
        if pats and not any(x in k for x in pats):
            continue
        t = m["gpu__time_duration.sum"] / 1e3
        rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
        print("%4s %-70s %8.1f us  rd %7.1f MB  wr %7.1f MB  %6.0f GB/s" % (i, k, t, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e3))


if __name__ == "__main__":
    main()
