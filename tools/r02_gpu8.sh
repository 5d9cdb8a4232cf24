#!/bin/bash
# Round-2 GPU check 3 (eight GPUs): bench at N=8 and N=4 (weak + strong legs) with NVLink byte counters around each.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
for n in 8 4; do
  nvidia-smi nvlink -gt d > gpurun_out/r02_nvlink_before_n$n.txt 2>&1
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 16 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo "bench$n rc=$?"; tail -4 gpurun_out/r02_bench_n$n.err
  nvidia-smi nvlink -gt d > gpurun_out/r02_nvlink_after_n$n.txt 2>&1
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02_bench_n%s.json" % n).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "per_rank_ms", "per_rank_sm_mhz")})
    s = d["strong"]
    print({k: s[k] for k in ("ms_per_frame", "value", "speedup_vs_1gpu", "efficiency_vs_1gpu", "bit_identical", "frames_checked", "per_rank_ms", "per_rank_sm_mhz")})
except Exception as e:
    print("no bench line", e)
PY
done
