"""Fold compute-sanitizer racecheck output (stdin) into counts per (hazard kind, kernel, reader line, writer line)."""
import collections
import re
import sys

counts = collections.Counter()
other = []
kind = None
threads = []
pending = []       # raw lines of the record being read
unmatched = []     # records that did not yield two thread lines
headers = 0
for line in sys.stdin:
    m = re.search(r"(?:Potential )?(\w+) hazard detected at (__shared__|__global__|\w+)", line)
    if m:
        if kind is not None and len(unmatched) < 5:
            unmatched.append(pending)
        headers += 1
        kind, threads, pending = m.group(1) + " " + m.group(2), [], [line.rstrip()]
        continue
    if kind is not None:
        pending.append(line.rstrip())
    m = re.search(r"(Read|Write) Thread .* at (.*?)\+0x[0-9a-f]+(?: in (\S+))?", line)
    if m and kind:
        kern = re.sub(r"\(.*", "", m.group(2)).split("::")[-1]
        threads.append("%s %s @ %s" % (m.group(1), kern, m.group(3) or "?"))
        if len(threads) == 2:
            counts[(kind, threads[0], threads[1])] += 1
            kind = None
        continue
    if "=========" not in line or "RACECHECK SUMMARY" in line or "ERROR SUMMARY" in line:
        other.append(line.rstrip())
print("racecheck hazards folded by (kind, first access, second access):")
for (k, a, b), n in counts.most_common():
    print("%9d  %-16s %s  |  %s" % (n, k, a, b))
print("total hazards folded: %d of %d hazard records" % (sum(counts.values()), headers))
for rec in unmatched:
    print("---- record without two thread lines:")
    print("\n".join(rec[:8]))
print("---- other output")
print("\n".join(other[-12:]))
