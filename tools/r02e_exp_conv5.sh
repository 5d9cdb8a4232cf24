#!/bin/bash
# round 2e: where do conv5's 137 cycles per MMA go?  Experiments build (wrong results), ncu time of the pair kernel.
mkdir -p gpurun_out
: > gpurun_out/r02e_exp_conv5.log
for v in 0 2 3; do
  INNFER_ROWS_DX0=$v ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_rows -s 0 -c 25 --csv --log-file gpurun_out/r02e_exp_conv5_$v.csv python tests/gpu_bringup.py --stage prof > /dev/null 2>&1
  echo "== INNFER_ROWS_DX0=$v" >> gpurun_out/r02e_exp_conv5.log
  python tools/ncu_seq.py gpurun_out/r02e_exp_conv5_$v.csv 0 0 >> gpurun_out/r02e_exp_conv5.log
done
cat gpurun_out/r02e_exp_conv5.log
