#!/bin/bash
# ncu passes on the GPU box: (1) launch list with device times, (2) full capture of conv kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches.csv python tests/gpu_bringup.py --stage prof > gpurun_out/prof1.log 2>&1
tail -3 gpurun_out/prof1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -s ${1:-40} -c ${2:-6} \
    -f -o gpurun_out/prof_conv python tests/gpu_bringup.py --stage prof > gpurun_out/prof2.log 2>&1
tail -3 gpurun_out/prof2.log
ls -la gpurun_out/
