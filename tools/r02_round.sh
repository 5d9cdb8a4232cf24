#!/bin/bash
# Round-2 record run (one gpurun call): whole GPU suite, smoke, the four bench workloads, reference arm
mkdir -p gpurun_out
T=${TAG:-r02d}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 > gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 400 gpurun_out/${T}_bench_n1.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/${T}_bench_reference_arm.json; echo
timeout 600 python bench.py --workload chain --steps 8 > gpurun_out/${T}_bench_chain.json 2>/dev/null; tail -c 200 gpurun_out/${T}_bench_chain.json; echo
timeout 600 python bench.py --workload small --steps 5 > gpurun_out/${T}_bench_small.json 2>/dev/null; tail -c 200 gpurun_out/${T}_bench_small.json; echo
timeout 600 python bench.py --workload i2i --steps 20 > gpurun_out/${T}_bench_i2i.json 2>/dev/null; tail -c 200 gpurun_out/${T}_bench_i2i.json; echo
for st in pan_time ppon_time srres_time; do timeout 300 python tests/gpu_bringup.py --stage $st 2>&1 | grep "iter=2" ; done
