#!/bin/bash
# round 2e: ncu --set full of one PPON _ResBlock_32 (c1, d1..d8, c2) and of the first RDB of the RRDB net
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 21 -c 11 -f -o gpurun_out/r02e_ppon_full \
    python tests/gpu_bringup.py --stage ppon_prof > gpurun_out/r02e_ppon_full.log 2>&1
tail -n 1 gpurun_out/r02e_ppon_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 16 -c 5 -f -o gpurun_out/r02e_rrdb_full \
    python tests/gpu_bringup.py --stage prof > gpurun_out/r02e_rrdb_full.log 2>&1
tail -n 1 gpurun_out/r02e_rrdb_full.log
ls -la gpurun_out/*.ncu-rep
