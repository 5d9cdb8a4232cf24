#!/bin/bash
# Round-2 experiment 1 (one gpurun call):
#  (a) frame time vs tile-batch size (does an L2-sized batch pay?),
#  (b) ncu DRAM bytes + duration of one RDB's conv launches at an L2-sized batch (4 tiles) and at 63 tiles,
#      caches NOT flushed between kernels (--cache-control none) so that L2 residency between launches is visible,
#  (c) compute-sanitizer racecheck over the conv-level GPU tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_smi.txt
INNFER_MB=95,48,24,12,8,6,5,4,3,2 timeout 300 python tests/gpu_bringup.py --stage time > gpurun_out/r02_mb_sweep.log 2>&1
tail -32 gpurun_out/r02_mb_sweep.log
for mb in 4 63; do
  INNFER_MB_PROF=$mb timeout 600 ncu --cache-control none --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    -k regex:conv_rows -s 200 -c 20 --csv --log-file gpurun_out/r02_l2_mb$mb.csv python tests/gpu_bringup.py --stage prof > gpurun_out/r02_l2_mb$mb.log 2>&1
  python - gpurun_out/r02_l2_mb$mb.csv $mb <<'PY'
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
by = {}
for row in csv.DictReader(lines):
    d = by.setdefault(row['ID'], {'k': row['Kernel Name'].split('conv_rows_kernel')[-1][:18]})
    d[row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
print("max_batch", sys.argv[2])
for v in by.values():
    print("%-20s %8.1f us  rd %8.1f MB  wr %8.1f MB  L2 hit %5.1f%%  pipe %4.1f%%" % (
        v['k'], v.get('gpu__time_duration.sum', 0) / 1e3, v.get('dram__bytes_read.sum', 0) / 1e6,
        v.get('dram__bytes_write.sum', 0) / 1e6, v.get('lts__t_sector_hit_rate.pct', 0),
        v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)))
PY
done
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv_block_wide_layout or conv_block_tcgen05 or rrdb_dense_blocks_with_amplified_weights" \
  > gpurun_out/r02_racecheck.log 2>&1
tail -15 gpurun_out/r02_racecheck.log
