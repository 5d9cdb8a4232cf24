#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -s 87 -c 6 -f -o gpurun_out/r02e_pan_tail_full python tests/gpu_bringup.py --stage pan_prof > gpurun_out/r02e_pan_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 12 -c 5 -f -o gpurun_out/r02e_pan_scpa_full python tests/gpu_bringup.py --stage pan_prof >> gpurun_out/r02e_pan_full.log 2>&1
tail -n 2 gpurun_out/r02e_pan_full.log
