#!/bin/bash
# round 2g: ncu --set full of one whole RRDB (15 conv launches) and of the tail with the final kernels ("after" rows)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 16 -c 15 -f -o gpurun_out/r02g_rrdb_full \
    python tests/gpu_bringup.py --stage prof > gpurun_out/r02g_rrdb_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 346 -c 6 -f -o gpurun_out/r02g_tail_full \
    python tests/gpu_bringup.py --stage prof >> gpurun_out/r02g_rrdb_full.log 2>&1
ls -la gpurun_out/r02g_*.ncu-rep
