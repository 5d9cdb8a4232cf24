#!/bin/bash
# DRAM / L2 view of 6 consecutive conv launches; usage: tools/gpu_exp2.sh "ENV=.." ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.max \
    --clock-control none ${CACHE:+--cache-control none} -k regex:conv_ -s ${SKIP:-40} -c 6 --csv --log-file gpurun_out/exp2_$i.csv python tests/gpu_bringup.py --stage prof > gpurun_out/exp2_$i.log 2>&1
  python - "$cfg" gpurun_out/exp2_$i.csv <<'PY'
import csv, sys
lines=[l for l in open(sys.argv[2]) if not l.startswith('==')]
by={}
for row in csv.DictReader(lines):
    d=by.setdefault(row['ID'],{'k':row['Kernel Name'].split('::')[-1][:12]})
    d[row['Metric Name']]=(float(row['Metric Value'].replace(',','')), row['Metric Unit'])
print("[%s]" % sys.argv[1])
for v in by.values():
    g=lambda k: v.get(k,(0,''))
    t=g('gpu__time_duration.sum')[0]/1e3
    def mb(k):
        x,u=g(k); return x*{'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}.get(u,1e-6)
    rd,wr=mb('dram__bytes_read.sum'),mb('dram__bytes_write.sum')
    print("  %-12s %7.1fus rd %6.0fMB wr %6.0fMB -> %4.2f TB/s dram%%=%.0f l2hit=%.0f tensor=%.0f kc=%.0f" % (v['k'], t, rd, wr, (rd+wr)/t/1e6*1e0, g('dram__throughput.avg.pct_of_peak_sustained_elapsed')[0], g('lts__t_sector_hit_rate.pct')[0], g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[0], g('sm__cycles_elapsed.max')[0]/1e3))
PY
done
