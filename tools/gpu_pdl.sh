#!/bin/bash
# A/B of programmatic dependent launch (INNFER_PDL) at small tile batches, full 1080p frame, alternating runs
mkdir -p gpurun_out
for cfg in "INNFER_PDL=0 INNFER_MB=5,2,1" "INNFER_PDL=1 INNFER_MB=5,2,1" "INNFER_PDL=0 INNFER_MB=5,2,1" "INNFER_PDL=1 INNFER_MB=5,2,1"; do
  echo "== $cfg"
  env $cfg timeout 300 python tests/gpu_bringup.py --stage time 2>&1 | grep "iter=2"
done
