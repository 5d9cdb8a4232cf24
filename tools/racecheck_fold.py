"""Fold compute-sanitizer racecheck output (stdin) into counts per (hazard kind, kernel, reader line, writer line)."""
import collections
import re
import sys

counts = collections.Counter()
other = []
kind = None
threads = []
for line in sys.stdin:
    m = re.search(r"Potential (\w+) hazard detected at (__shared__|__global__|\w+)", line)
    if m:
        kind, threads = m.group(1) + " " + m.group(2), []
        continue
    m = re.search(r"(Read|Write) Thread .* at (.*?)\+0x[0-9a-f]+ in (\S+)", line)
    if m and kind:
        kern = re.sub(r"\(.*", "", m.group(2)).split("::")[-1]
        threads.append("%s %s @ %s" % (m.group(1), kern, m.group(3)))
        if len(threads) == 2:
            counts[(kind, threads[0], threads[1])] += 1
            kind = None
        continue
    if "=========" not in line or "RACECHECK SUMMARY" in line or "ERROR SUMMARY" in line:
        other.append(line.rstrip())
print("racecheck hazards folded by (kind, first access, second access):")
for (k, a, b), n in counts.most_common():
    print("%9d  %-16s %s  |  %s" % (n, k, a, b))
print("total hazards: %d" % sum(counts.values()))
print("---- other output")
print("\n".join(other[-12:]))
