#!/bin/bash
# round 2f, two GPUs: 2-GPU bit-identity test + bench at N=2 (weak + strong legs) with the round-2e/f kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02f_pytest_multi.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/r02f_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-torch-gpu --no-cpu-baseline > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench2 rc=$?"; tail -4 gpurun_out/r02f_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "per_rank_ms", "per_rank_sm_mhz")})
    print(json.dumps(d["strong"])[:1200])
except Exception as e:
    print("no bench line", e)
PY
