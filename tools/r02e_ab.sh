#!/bin/bash
# round 2e A/B on one box: 16-warp epilogue of the Cout = 32 row kernels (INNFER_ROWS_WEPI), alternating runs
mkdir -p gpurun_out
: > gpurun_out/r02e_ab.log
for rep in 1 2; do
  for v in 0 3; do
    echo "== INNFER_ROWS_WEPI=$v" >> gpurun_out/r02e_ab.log
    INNFER_ROWS_WEPI=$v INNFER_MB=95 python tests/gpu_bringup.py --stage time 2>&1 | grep "time 1080p" >> gpurun_out/r02e_ab.log
  done
done
echo "== PPON" >> gpurun_out/r02e_ab.log
python tests/gpu_bringup.py --stage ppon_time 2>&1 | grep "ppon 1080p" >> gpurun_out/r02e_ab.log
cat gpurun_out/r02e_ab.log
