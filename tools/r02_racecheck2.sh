#!/bin/bash
# racecheck over every conv-level GPU test incl. both image-to-image kernels (im2col and halo-tile)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10000000 --show-backtrace no \
  python -m pytest tests/test_gpu_parity.py tests/test_i2i.py -q -m gpu \
  -k "conv_block_wide_layout or conv_block_tcgen05 or rrdb_dense_blocks_with_amplified_weights or generator_layer_tcgen05" 2>&1 \
  | python tools/racecheck_fold.py > gpurun_out/r02c_racecheck_summary.txt
tail -25 gpurun_out/r02c_racecheck_summary.txt
