"""Regenerates profiles/sass_summary.txt from the built library: per kernel, counts of the SASS mnemonics that prove the
Blackwell path (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, .WS = weight stationary; LDTM = tcgen05.ld; UTMALDG = TMA
tensor load; UBLKCP = cp.async.bulk; LDGSTS = cp.async; UTCBAR = tcgen05.commit; SYNCS = mbarrier; FFMA for the fp32 kernels).

    python tools/sass_summary.py [round tag]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "innfer_b200", "lib", "libinnfer_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "round 2e"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
want = re.compile(r"\b(UTCHMMA(?:\.2CTA|\.WS)*|LDTM(?:\.x\d+)?|UTMALDG(?:\.\dD)?|UBLKCP(?:\.S\.G)?|LDGSTS|UTCBAR|UTCATOMSWS|SYNCS|FFMA|STTM)\b")
out = ["# SASS summary of innfer_b200/lib/libinnfer_b200.so (cuobjdump -sass, sm_100a), %s; written by tools/sass_summary.py" % tag,
       "# per kernel: counts of tensor-core (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2), TMEM load (LDTM = tcgen05.ld), TMA tensor load",
       "# (UTMALDG), bulk copy (UBLKCP = cp.async.bulk), cp.async (LDGSTS), tcgen05.commit (UTCBAR), mbarrier (SYNCS) instructions; FFMA for the fp32-mode kernels",
       ""]
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    cnt = collections.Counter()
    for line in blk.split("\n"):
        m = want.search(line)
        if m and "/*" in line:
            key = m.group(1)
            if key.startswith("SYNCS"):
                key = "SYNCS"
            cnt[key] += 1
    short = name.replace("innfer::(anonymous namespace)::", "").replace("innfer::<unnamed>::", "")
    short = re.sub(r"\((int|bool)\)", "", short)          # cu++filt prints template arguments as (int)64, (bool)1
    short = short[:short.index(">(") + 1] if ">(" in short else short.split("(")[0]
    if any(k.startswith(("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS")) for k in cnt) or cnt.get("FFMA", 0) > 50:
        out.append("%-70s %s" % (short, "  ".join("%s=%d" % kv for kv in sorted(cnt.items()))))
tot = collections.Counter()
for l in out[4:]:
    for kv in l[70:].split():
        k, v = kv.split("=")
        tot[k] += int(v)
out.append("")
out.append("totals: " + "  ".join("%s=%d" % kv for kv in sorted(tot.items())))
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
print(out[-1])
