#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "rrdb or fixture or conv_block_wide or determinism or full_size" > gpurun_out/r02f_pytest_sub.log 2>&1; tail -2 gpurun_out/r02f_pytest_sub.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_rows -s 0 -c 45 --csv --log-file gpurun_out/r02f_c5.csv python tests/gpu_bringup.py --stage prof > /dev/null 2>&1
python tools/ncu_seq.py gpurun_out/r02f_c5.csv 0 0
INNFER_MB=95 python tests/gpu_bringup.py --stage time 2>&1 | grep "time 1080p"
