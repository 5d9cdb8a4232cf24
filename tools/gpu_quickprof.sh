#!/bin/bash
# quick ncu pass: a few pipe metrics on a handful of conv launches
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -k regex:conv_ -s ${1:-40} -c ${2:-6} --csv --log-file gpurun_out/quick.csv python tests/gpu_bringup.py --stage prof > gpurun_out/quick.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/quick.csv') if not l.startswith('==')]
r=list(csv.DictReader(lines))
by={}
for row in r:
    by.setdefault(row['ID'],{})[row['Metric Name']]=row['Metric Value']
    by[row['ID']]['k']=row['Kernel Name'][30:50]
for k,v in by.items(): print(k,v)
PY
