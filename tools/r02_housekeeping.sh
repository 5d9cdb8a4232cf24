#!/bin/bash
# Round-2 housekeeping on the GPU box (one gpurun call):
#  (a) ncu metrics of the trunk conv kernels at the BENCH's batch size (95 tiles of a 1080p frame): DRAM bytes,
#      duration, tensor-pipe activity per launch -> gpurun_out/r02_conv_metrics.csv (tools/conv_metrics.py turns it
#      into profiles/r02_conv_metrics.json, which bench.py quotes as `traffic` / per-family `ncu`)
#  (b) compute-sanitizer racecheck over every conv-level GPU test (RRDB path and image-to-image layers), all hazards
#      printed without backtraces and folded into (kind, reader line, writer line) counts
mkdir -p gpurun_out
timeout 900 ncu --cache-control none --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  -k regex:conv_ -s 61 -c 45 --csv --log-file gpurun_out/r02_conv_metrics.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu > gpurun_out/r02_conv_metrics.log 2>&1
tail -2 gpurun_out/r02_conv_metrics.log
python tools/conv_metrics.py gpurun_out/r02_conv_metrics.csv gpurun_out/r02_conv_metrics.json
timeout 1700 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10000000 --show-backtrace no \
  python -m pytest tests/test_gpu_parity.py tests/test_i2i.py -q -m gpu \
  -k "conv_block_wide_layout or conv_block_tcgen05 or rrdb_dense_blocks_with_amplified_weights or generator_layer_tcgen05" 2>&1 \
  | python tools/racecheck_fold.py > gpurun_out/r02_racecheck_summary.txt
cat gpurun_out/r02_racecheck_summary.txt | tail -30
