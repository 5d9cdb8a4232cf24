#!/bin/bash
# Round-2e housekeeping on the GPU box (one gpurun call):
#  (a) ncu metrics of the trunk conv kernels at the bench's batch size -> profiles/r02_conv_metrics.json (bench.py's `traffic`)
#  (b) compute-sanitizer racecheck + memcheck over the tests that cover the kernels changed in round 2e: 16-warp epilogues
#      (conv_rows / conv_tc / conv_up), identity-MMA residual, dilated kernels with register prefetch, tiled colour fix,
#      uint8 image_to_tiles with the pad-chunk skip
mkdir -p gpurun_out
timeout 900 ncu --cache-control none --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  -k regex:conv_ -s 61 -c 45 --csv --log-file gpurun_out/r02_conv_metrics.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu > gpurun_out/r02_conv_metrics.log 2>&1
tail -2 gpurun_out/r02_conv_metrics.log
python tools/conv_metrics.py gpurun_out/r02_conv_metrics.csv gpurun_out/r02_conv_metrics.json
SEL="conv_block_wide_layout or conv_block_tcgen05 or rrdb_dense_blocks_with_amplified_weights or ppon_vs_reference_fixture or pan_vs_reference_fixture or ppon_dilated_branch or color_fix or pad_chunk or pixel_kernels"
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10000000 --show-backtrace no \
  python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$SEL" 2>&1 | python tools/racecheck_fold.py > gpurun_out/r02e_racecheck_summary.txt
tail -22 gpurun_out/r02e_racecheck_summary.txt
timeout 1500 compute-sanitizer --tool memcheck --print-limit 100 \
  python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$SEL" > gpurun_out/r02e_memcheck.txt 2>&1
tail -6 gpurun_out/r02e_memcheck.txt
