#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "pan or batch_forward or small" > gpurun_out/r02f_pytest_pan.log 2>&1; tail -2 gpurun_out/r02f_pytest_pan.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02f_pan.csv python tests/gpu_bringup.py --stage pan_prof > /dev/null 2>&1
python tools/ncu_seq.py gpurun_out/r02f_pan.csv 83 5 | tail -18
python tests/gpu_bringup.py --stage pan_time 2>&1 | grep "fp16 iter"
