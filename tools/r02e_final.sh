#!/bin/bash
# round 2e final single-GPU run: whole GPU suite, smoke, family frame times (A/B of the Cout = 64 wide epilogue), launch
# list of the RRDB frame, the bench workloads and the reference arm
mkdir -p gpurun_out
T=${TAG:-r02f}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4 > gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in 3 15; do echo "== INNFER_ROWS_WEPI=$v"; for st in srres_time ppon_time; do INNFER_ROWS_WEPI=$v timeout 300 python tests/gpu_bringup.py --stage $st 2>&1 | grep "iter=2"; done; done
timeout 300 python tests/gpu_bringup.py --stage pan_time 2>&1 | grep "fp16 iter=2"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_rrdb.csv python tests/gpu_bringup.py --stage prof > /dev/null 2>&1
python tools/ncu_seq.py gpurun_out/${T}_rrdb.csv 340 14 | tail -26
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 300 gpurun_out/${T}_bench_n1.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>/dev/null
timeout 600 python bench.py --workload chain --steps 8 > gpurun_out/${T}_bench_chain.json 2>/dev/null; tail -c 200 gpurun_out/${T}_bench_chain.json; echo
timeout 600 python bench.py --workload small --steps 5 > gpurun_out/${T}_bench_small.json 2>/dev/null; tail -c 300 gpurun_out/${T}_bench_small.json; echo
