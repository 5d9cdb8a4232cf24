#!/bin/bash
# Round-2 record run (one gpurun call): whole GPU suite, i2i bench workload, i2i launch list + ncu captures of its kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r02b_pytest.log; tail -4 gpurun_out/r02b_pytest.log
timeout 600 python bench.py --workload i2i --steps 20 --warmup 3 > gpurun_out/r02b_bench_i2i.json 2> gpurun_out/r02b_bench_i2i.err; tail -c 600 gpurun_out/r02b_bench_i2i.json
STAGE=i2i_prof bash tools/gpu_launchlist.sh r02b_i2i 2>&1 | tail -9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gen_conv|norm_" -s 230 -c 12 -f -o gpurun_out/r02b_i2i_kernels python tests/gpu_bringup.py --stage i2i_prof > gpurun_out/r02b_i2i_ncu.log 2>&1; tail -2 gpurun_out/r02b_i2i_ncu.log
