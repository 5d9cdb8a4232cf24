#!/bin/bash
# round 2g: tile-batch-size sweep with the round-2g kernels (the dilated convs are HBM-bound now: do L2 hits pay?)
mkdir -p gpurun_out
{
echo "== PPON 1080p (ms per frame, iter 1 and 2)"
INNFER_MB=95,48,32,24,16,12,8 python tests/gpu_bringup.py --stage ppon_time 2>&1 | grep "iter=[12]"
echo "== RRDB 1080p"
INNFER_MB=95,64,48,32,24 python tests/gpu_bringup.py --stage time 2>&1 | grep "iter=[12]"
} > gpurun_out/r02g_batch_sweep.txt 2>&1
cat gpurun_out/r02g_batch_sweep.txt
