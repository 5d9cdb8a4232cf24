#!/bin/bash
# Runs the bring-up stages on the GPU box, each under its own timeout, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi.txt
for st in "$@"; do
  echo "=== stage $st"
  timeout 300 python tests/gpu_bringup.py --stage $st 2>&1 | tee gpurun_out/bringup_$st.log | tail -60
done
