#!/bin/bash
# round 2f, eight GPUs: bench at N=8 (weak + strong legs) with the final kernels
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 16 --warmup 3 --no-torch-gpu --no-cpu-baseline > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r02f_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench_n8.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "per_rank_ms", "per_rank_sm_mhz")})
    s = d["strong"]
    print({k: s[k] for k in ("ms_per_frame", "value", "speedup_vs_1gpu", "efficiency_vs_1gpu", "bit_identical", "frames_checked", "per_rank_ms", "per_rank_sm_mhz")})
except Exception as e:
    print("no bench line", e)
PY
