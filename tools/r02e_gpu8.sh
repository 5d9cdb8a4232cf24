#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest.log 2>&1; tail -3 gpurun_out/r02e_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02e_ppon.csv python tests/gpu_bringup.py --stage ppon_prof > gpurun_out/r02e_ppon.log 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02e_pan.csv python tests/gpu_bringup.py --stage pan_prof > gpurun_out/r02e_pan.log 2>&1
python tests/gpu_bringup.py --stage pan_time 2>&1 | grep "pan 1080p"
python tests/gpu_bringup.py --stage ppon_time 2>&1 | grep "ppon 1080p"
python tests/gpu_bringup.py --stage srres_time 2>&1 | grep "srresnet 1080p"
