#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest.log 2>&1; tail -3 gpurun_out/r02e_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02e_rrdb.csv python tests/gpu_bringup.py --stage prof > gpurun_out/r02e_rrdb.log 2>&1
tail -n 1 gpurun_out/r02e_rrdb.log
INNFER_MB=95 python tests/gpu_bringup.py --stage time 2>&1 | grep "time 1080p"
