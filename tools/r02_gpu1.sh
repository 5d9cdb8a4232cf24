#!/bin/bash
# Round-2 GPU check 1 (one GPU): full GPU parity suite, smoke, the default bench line and the chain workload.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_a.err; cut -c1-1500 gpurun_out/r02_bench_a.json
timeout 600 python bench.py --workload chain --steps 8 --warmup 3 > gpurun_out/r02_bench_chain.json 2> gpurun_out/r02_bench_chain.err; echo "chain rc=$?"; tail -5 gpurun_out/r02_bench_chain.err; cut -c1-1200 gpurun_out/r02_bench_chain.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02_bench_ref.json
