#!/bin/bash
# ncu launch list (device time per launch) of one 800x1000 frame; usage: tools/gpu_launchlist.sh TAG [ENV=..]
mkdir -p gpurun_out
tag=$1; shift
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_$tag.csv python tests/gpu_bringup.py --stage ${STAGE:-prof} > gpurun_out/prof_$tag.log 2>&1
tail -2 gpurun_out/prof_$tag.log
python - gpurun_out/launches_$tag.csv <<'PY'
import csv, sys, collections
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
tot=collections.OrderedDict()
seq=[]
for r in rows:
    k=r['Kernel Name'].split('::')[-1].split('(')[0]
    v=float(r['Metric Value'].replace(',',''))/1e3
    d=tot.setdefault(k,[0,0.0]); d[0]+=1; d[1]+=v
    seq.append((k,v))
T=sum(v[1] for v in tot.values())
for k,(n,v) in tot.items(): print("%-40s n=%4d total=%9.1f us avg=%8.1f share=%.1f%%"%(k,n,v,v/n,100*v/T))
print("total %.1f ms"%(T/1e3))
print("first RRDB:", " ".join("%.0f"%v for k,v in seq[1:18]))
print("tail:", " ".join("%s:%.0f"%(k[:9],v) for k,v in seq[-8:]))
PY
