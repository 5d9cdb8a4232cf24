#!/bin/bash
# Timing experiments: ncu durations + pipe metrics of 6 consecutive conv launches (one RDB) per env setting.
# usage: tools/gpu_exp.sh "ENV1=a ENV2=b" "ENV3=c" ...   (one quoted env string per configuration; "" = defaults)
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max \
    --clock-control none -k regex:conv_ -s ${SKIP:-40} -c ${COUNT:-6} --csv --log-file gpurun_out/exp_$i.csv python tests/gpu_bringup.py --stage prof > gpurun_out/exp_$i.log 2>&1
  python - "$cfg" gpurun_out/exp_$i.csv <<'PY'
import csv, sys
lines=[l for l in open(sys.argv[2]) if not l.startswith('==')]
by={}
for row in csv.DictReader(lines):
    d=by.setdefault(row['ID'],{'k':row['Kernel Name'].split('::')[-1][:14]})
    d[row['Metric Name']]=float(row['Metric Value'].replace(',',''))
out=[]
for v in by.values():
    out.append("%s %.1fus %.0fkc %.0f%%" % (v['k'], v.get('gpu__time_duration.sum',0), v.get('sm__cycles_elapsed.max',0)/1e3, v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',0)))
print("[%s]" % sys.argv[1], " | ".join(out))
PY
done
