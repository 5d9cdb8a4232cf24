#!/bin/bash
# kernel durations with warm L2 (no cache flush between launches) for several tile-batch sizes
mkdir -p gpurun_out
for mb in 4 6 8 12 38; do
  INNFER_MB_PROF=$mb timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
    -k regex:conv_ -s 41 -c 5 --csv --log-file gpurun_out/cache_$mb.csv python tests/gpu_bringup.py --stage prof > gpurun_out/cache.log 2>&1
  python - <<PY
import csv
lines=[l for l in open('gpurun_out/cache_$mb.csv') if not l.startswith('==')]
by={}
for row in csv.DictReader(lines):
    by.setdefault(row['ID'],{})[row['Metric Name']]=float(row['Metric Value'].replace(',',''))
print('mb=$mb', [(round(v['gpu__time_duration.sum']/1e3,1), round(v['dram__bytes_read.sum']/1e6), round(v['dram__bytes_write.sum']/1e6)) for v in by.values()])
PY
done
