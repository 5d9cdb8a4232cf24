"""Headline benchmark: output Mpix/s of 4x ESRGAN RRDB (23 blocks, nf=64) fp16 on synthetic
1920x1080 frames with chop_forward tiling (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One step = one 1080p frame through the hot path (tile -> 351 fused convs per tile batch -> blend
-> uint8).  N > 1: every rank upscales its own frame (the path shards by image, no collective on
the data path); torch.distributed is used only for the barrier and the max-over-ranks time.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, SCALE, NB, NF = 1080, 1920, 4, 23, 64
PATCH, STEP = 200, 0.5
WORKLOAD = "4x ESRGAN RRDBNet (23 RRDB, nf=64, random-init seed 0) fp16, synthetic 1920x1080 frames, chop_forward 200px tiles step 0.5 (190 tiles/frame)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernels are timed inside a long step)"
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def synth_frame(seed):
    return np.random.default_rng(seed).integers(0, 256, (H, W, 3), dtype=np.uint8)


def cpu_reference_rate(n_tiles, threads=None):
    """Times the oracle (CPU restatement of the reference forward, fp32, torch CPU) on the first
    n_tiles 200x200 tiles of frame 0; returns (output Mpix/s extrapolated to a frame, seconds, threads)."""
    from oracle import rrdb_oracle as O
    if threads is None:
        # torchrun exports OMP_NUM_THREADS=1; the CPU legs use every core this process may run on
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(max(1, threads))
    sd = O.make_state_dict(scale=SCALE, nb=NB, nf=NF, seed=0)
    x = O.np2tensor(synth_frame(0))
    patches, _, _ = O.extract_patches(x, PATCH, STEP)
    O.rrdbnet_forward(sd, patches[0:1, :, :32, :32], SCALE)  # warm-up (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    for i in range(n_tiles):
        O.rrdbnet_forward(sd, patches[i:i + 1], SCALE)
    dt = time.perf_counter() - t0
    frame_seconds = dt / n_tiles * patches.shape[0]
    return (SCALE * H * SCALE * W) / frame_seconds / 1e6, dt, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    tiles = max(1, args.ref_tiles)
    rates = []
    for _ in range(args.warmup):
        pass  # the oracle warms itself up inside cpu_reference_rate
    total = 0.0
    for _ in range(args.steps):
        r, dt, threads = cpu_reference_rate(tiles)
        rates.append(r)
        total += dt
    value = float(np.mean(rates))
    sample = "%d of 190 tiles (200x200, 4x net, fp32) per step, extrapolated x190/%d to a frame" % (tiles, tiles)
    line = {
        "impl": "reference", "metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_tile_sharded(args, eng, dist, rank, world, local, warm):
    """Strong scaling on single frames: every rank computes ceil(190/N) tiles of each frame and its
    last conv stores them into the owner's tile buffer over NVLink (innfer_b200/multi_gpu.py)."""
    from innfer_b200 import multi_gpu as MG
    be = MG.NativeTileBackend(eng, H, W, PATCH, STEP)
    up = MG.TileShardedUpscaler(be, dist)
    frames = [synth_frame(i) for i in range(2)]
    out_pix = SCALE * H * SCALE * W
    f = 0
    for _ in range(warm):
        up.upscale(f, frames[f % 2] if MG.frame_owner(f, world) == rank else None)
        f += 1
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        up.upscale(f, frames[f % 2] if MG.frame_owner(f, world) == rank else None)
        f += 1
    torch.cuda.synchronize()
    dist.barrier()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    up.close()
    if rank == 0:
        line = {"metric": "output Mpix/s", "value": out_pix / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
                "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "sharding": "tiles of one frame across ranks, CUDA-IPC peer stores, "
                                                             "owner-side blend, 2 host barriers per frame",
                           "timing": "host wall clock around whole frames incl. H2D of the frame and D2H of the result"},
                "e2e": {"value": out_pix / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": ms,
                        "h2d_bytes_per_step": H * W * 3, "d2h_bytes_per_step": out_pix * 3}}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-tiles", type=int, default=2, help="tiles per step for the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-batch", type=int, default=0)
    ap.add_argument("--shard", default="images", choices=["images", "tiles"],
                    help="N>1: 'images' = one frame per rank per step (weak scaling, default); 'tiles' = all ranks "
                         "split the tiles of ONE frame and stitch through CUDA-IPC peer stores (strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU leg)")
    warm = max(3, args.warmup)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from innfer_b200 import _native as N
    from innfer_b200 import synth
    from innfer_b200.engine import RRDBEngine

    sd = synth.make_state_dict(scale=SCALE, nb=NB, nf=NF, seed=0)
    cfg = dict(in_nc=3, out_nc=3, nf=NF, nb=NB, gc=32, scale=SCALE, plus=False)
    eng = RRDBEngine.from_state_dict(sd, cfg, dev, fp16=True)
    if args.max_batch:
        eng.set_max_batch(args.max_batch)

    if args.shard == "tiles" and world > 1:
        run_tile_sharded(args, eng, dist, rank, world, local, warm)
        return

    frames = [synth_frame(1000 * rank + i) for i in range(2)]
    d_in = [torch.from_numpy(f).to(dev) for f in frames]
    d_out = torch.empty((SCALE * H, SCALE * W, 3), dtype=torch.uint8, device=dev)
    h_in = [torch.from_numpy(f).pin_memory() for f in frames]
    h_out = torch.empty((SCALE * H, SCALE * W, 3), dtype=torch.uint8).pin_memory()
    out_pix = SCALE * H * SCALE * W

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident leg ("value")
    for i in range(warm):
        eng.upscale_u8_device(d_in[i % 2], PATCH, STEP, out=d_out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile_reset()
    launches0 = N.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.upscale_u8_device(d_in[i % 2], PATCH, STEP, out=d_out)
    e1.record()
    barrier()
    launches = N.kernel_launches() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    conv_ms, conv_launches = eng.profile_read()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * out_pix / (ms_per_step * 1e-3) / 1e6

    # ---------------- end-to-end leg: pinned host uint8 in, pinned host uint8 out, copies timed
    for i in range(2):
        eng.upscale_u8(h_in[i % 2], PATCH, STEP, out=h_out)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        eng.upscale_u8(h_in[i % 2], PATCH, STEP, out=h_out)
    t1.record()
    barrier()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1)) / args.steps
    e2e_value = world * out_pix / (e2e_ms * 1e-3) / 1e6
    checksum = int(h_out[::97, ::89].to(torch.int64).sum().item())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    flop_step = 190 * PATCH * PATCH * synth.flop_per_lr_pixel(SCALE, NB, NF)
    roof = None
    if conv_launches:
        achieved = flop_step * args.steps / (conv_ms * 1e-3) / 1e12
        peak = float(peaks["bf16_tflops_sustained"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch_avg")
        # the same conv sequence seen from the HBM side: conv1..conv4 of every dense block (row-streaming
        # kernel, 190-240 FLOP per byte) run at the DRAM roofline when timed alone (profiles/)
        bytes_step = 190 * PATCH * PATCH * synth.bytes_per_lr_pixel(SCALE, NB, NF)
        hbm_peak = float(peaks["hbm_gbs"])
        hbm_achieved = bytes_step * args.steps / (conv_ms * 1e-3) / 1e9
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": "conv_rows_kernel<32|64,K,RES,PAIR> (+ conv_up / conv_tc for 6 tail convs): all %d conv launches of a step" % (conv_launches // args.steps),
                "flop_per_launch_avg": flop_step * args.steps / conv_launches,
                "avg_launch_ms": conv_ms / conv_launches, "peak_source": peak_src,
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "algorithmic_bytes_per_step": bytes_step}}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r, dt, threads = cpu_reference_rate(args.ref_tiles)
        cpu = {"value": r, "unit": "Mpix/s", "cores": threads, "kind": "port",
               "sample": "%d of 190 tiles of one frame through oracle/rrdb_oracle.py (torch CPU fp32, %.1f s), extrapolated to a frame"
                         % (args.ref_tiles, dt)}
    line = {
        "metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "sharding": "by image, no collective",
                   "l2": "per-step working set (GBs of tile activations) >> 126 MB L2, no flush needed",
                   "checksum": checksum},
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": H * W * 3,
                "d2h_bytes_per_step": out_pix * 3},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
