#!/bin/bash
# round 2e: full GPU suite with the 16-warp epilogue default, then the default bench line
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest.log 2>&1; tail -3 gpurun_out/r02e_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; tail -c 600 gpurun_out/r02e_bench_n1.json
INNFER_ROWS_WEPI=0 python bench.py --steps 8 --warmup 3 > gpurun_out/r02e_bench_n1_wepi0.json 2> gpurun_out/r02e_bench_n1_wepi0.err
