"""Headline benchmark: output Mpix/s of 4x ESRGAN RRDB (23 blocks, nf=64) fp16 on synthetic
1920x1080 frames with chop_forward tiling (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch-gpu] [--workload frame|chain]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One step = one frame through the hot path (tile -> 351 fused convs per tile batch -> blend -> uint8).  Prints ONE
JSON line on rank 0.

N = 1: `value` (device-resident frames), `e2e` (pinned host uint8 in -> pinned host uint8 out through the C-ABI call,
copies inside the timed region), `roofline` (aggregate + one entry per conv kernel family), `cpu_baseline` (the
reference's CPU path on the host cores) and `torch_gpu_baseline` (the unmodified reference on the same GPU through
PyTorch/cuDNN fp16, its own per-tile loop -- the "kernel to beat").
N > 1: the weak-scaling leg (every rank upscales its own frames, no exchange) gives `value`; the strong-scaling leg
(`strong`: all ranks split the tiles of every single frame, CUDA-IPC peer stores over NVLink, device-flag pipeline,
innfer_b200/multi_gpu.py) gives ms per frame, efficiency against this run's own 1-GPU frame time and a bit-identity
check of the stitched frames against rank 0's single-GPU result.  No collective on the data path; torch.distributed
carries the barrier, the max-over-ranks reductions and small Python objects.

--workload chain: BASELINE configs[2], 1x RRDB (JPEG denoise) + 4x RRDB with -cf on 1280x720 frames through the
same device pipeline run.py uses (innfer_b200.run.ChainRunner).
"""
import argparse
import contextlib
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCALE, NB, NF = 4, 23, 64
PATCH, STEP = 200, 0.5
WORKLOADS = {
    "frame": {"H": 1080, "W": 1920, "chain": [4], "cf": False, "tiles": 190,
              "name": "4x ESRGAN RRDBNet (23 RRDB, nf=64, random-init seed 0) fp16, synthetic 1920x1080 frames, "
                      "chop_forward 200px tiles step 0.5 (190 tiles/frame)"},
    "chain": {"H": 720, "W": 1280, "chain": [1, 4], "cf": True, "tiles": 84,
              "name": "chained 1x RRDB (23 blocks) + 4x RRDB (23 blocks) fp16 with -cf colour fix, synthetic 1280x720 "
                      "frames, chop_forward 200px tiles step 0.5 (84 + 84 tiles/frame)"},
    "small": {"name": "4x SRResNet (16 blocks) / PAN (16 SCPA, self attention) / PPON (24 blocks) at their get_network_G_config "
                      "defaults, random-init, fp16, synthetic 512x512 frames, chop_forward 200px tiles step 0.5 (25 tiles/frame); "
                      "headline = SRResNet"},
    "i2i": {"name": "pix2pix unet_256 (train-mode BatchNorm, whole image) and CycleGAN resnet_9blocks (InstanceNorm), ngf 64, "
                    "random-init, fp16, synthetic 256x256 and 1024x1024 images, batch 1; headline = resnet_9blocks at 1024x1024"},
}
H, W = WORKLOADS["frame"]["H"], WORKLOADS["frame"]["W"]
WORKLOAD = WORKLOADS["frame"]["name"]
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernels are timed inside a long step)"
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while a timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, cuda_index):
        self.rows, self.proc = [], None
        self.sel = str(cuda_index)
        try:  # nvidia-smi numbers GPUs independently of CUDA_VISIBLE_DEVICES: select by UUID when torch exposes it
            u = str(torch.cuda.get_device_properties(cuda_index).uuid)
            self.sel = u if u.startswith("GPU-") else "GPU-" + u
        except Exception:
            pass

    def start(self, settle=0.0):
        """Starts sampling; `settle` seconds are slept so that nvidia-smi is up before a short timed region begins
        (samples are filtered to the [mark_begin(), stop()] window by their arrival time)."""
        self.rows, self.t_begin = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.sel, "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, args=(self.proc,), daemon=True).start()
        except OSError:
            self.proc = None
        if settle:
            time.sleep(settle)

    def mark_begin(self):
        self.t_begin = time.time()

    def _read(self, proc):
        for line in proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts + [time.time()])

    def stop(self):
        t_end = time.time()
        if self.proc is not None:
            self.proc.terminate()
            self.proc = None
        rows = list(self.rows)
        if self.t_begin is not None:
            rows = [r for r in rows if self.t_begin <= r[-1] <= t_end + 0.05]
        sm = sorted(int(r[0]) for r in rows if r[0].isdigit())
        mx = [int(r[1]) for r in rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


def synth_frame(seed, h=H, w=W):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def host_threads():
    # torchrun exports OMP_NUM_THREADS=1; the CPU legs use every core this process may run on
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ reference arms
def load_reference():
    """The unmodified reference (a git-ignored copy of /root/reference under baseline/_ref, made by
    __graft_entry__.build() where the mount exists; it travels to the GPU box with the snapshot), imported under its
    own top-level module names (run, utils, architectures).  None when the copy is absent."""
    if not os.path.isfile(os.path.join(REF_DIR, "run.py")):
        return None
    import importlib.util
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    spec = importlib.util.spec_from_file_location("innfer_reference_run", os.path.join(REF_DIR, "run.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def save_synth_model(dirname, scale=SCALE):
    from innfer_b200 import synth
    path = os.path.join(dirname, "%dx_rand_rrdb.pth" % scale)
    torch.save(synth.make_state_dict(scale=scale, nb=NB, nf=NF, seed=0), path)
    return path


class CpuReference:
    """The reference's CPU path (run.py -cpu: fp32, its own Model / chop_forward / per-tile loop) on a bounded sample:
    a 200 x (100 * (n + 1)) crop of frame 0 = n tiles of 200x200 through Model.__call__, extrapolated x190/n to a
    frame.  kind "reference" when baseline/_ref is present, else the oracle port of the same functions ("port")."""

    def __init__(self, tiles):
        self.tiles = max(1, int(tiles))
        self.threads = host_threads()
        torch.set_num_threads(self.threads)
        self.img = synth_frame(0)[:PATCH, :PATCH // 2 * (self.tiles + 1)]
        with contextlib.redirect_stdout(sys.stderr):   # the reference prints; stdout carries the one JSON line
            ref = load_reference()
        if ref is not None:
            self.kind = "reference"
            self.td = tempfile.TemporaryDirectory()
            with contextlib.redirect_stdout(sys.stderr):
                self.model = ref.Model(save_synth_model(self.td.name), "infer", None, device=torch.device("cpu"))
            from utils.utils import np2tensor, tensor2np   # the reference's own (baseline/_ref is first on sys.path)
            self.np2tensor, self.tensor2np = np2tensor, tensor2np
        else:
            from oracle import rrdb_oracle as O
            self.kind = "port"
            self.O = O
            self.sd = O.make_state_dict(scale=SCALE, nb=NB, nf=NF, seed=0)

    def step(self):
        """One pass over the sample; returns seconds."""
        t0 = time.perf_counter()
        if self.kind == "reference":
            with contextlib.redirect_stdout(sys.stderr):
                out = self.tensor2np(self.model(self.np2tensor(self.img)).detach())
        else:
            out = self.O.tensor2np(self.O.chop_forward(self.sd, self.O.np2tensor(self.img), patch_size=PATCH))
        assert out.shape == (SCALE * self.img.shape[0], SCALE * self.img.shape[1], 3)
        return time.perf_counter() - t0

    def rate(self, seconds):
        """output Mpix/s of a whole 1080p frame extrapolated from `seconds` per sample"""
        return (SCALE * H * SCALE * W) / (seconds / self.tiles * WORKLOADS["frame"]["tiles"]) / 1e6

    def sample(self):
        return ("%d of 190 tiles (a 200x%d crop of frame 0 through %s, fp32, %d threads), extrapolated x190/%d to a frame"
                % (self.tiles, self.img.shape[1], "the reference's run.Model.__call__ (chop_forward, per-tile loop, "
                   "recompose_tensor)" if self.kind == "reference" else "oracle/rrdb_oracle.py chop_forward", self.threads,
                   self.tiles))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cpu = CpuReference(args.ref_tiles)
    for _ in range(max(1, min(args.warmup, 2))):   # thread pool, oneDNN primitive caches (each pass is seconds of CPU work)
        cpu.step()
    times = [cpu.step() for _ in range(args.steps)]
    sec = float(np.mean(times))
    value = cpu.rate(sec)
    line = {
        "impl": "reference", "metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "sample": cpu.sample()},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": cpu.threads, "kind": cpu.kind, "sample": cpu.sample()},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_torch_gpu(args, stdout=sys.stdout):
    """--impl torch-gpu: the unmodified reference on the GPU the way its main() runs it (run.py:336-337,382-431:
    cudnn.benchmark, default tensor type cuda.HalfTensor, Model.__call__ with its per-tile loop, np2tensor / tensor2np
    around it) on whole 1080p frames; falls back to the oracle's torch functions on CUDA fp16 when baseline/_ref is
    absent.  This is PyTorch-dispatched cuDNN on the same B200: the kernel-quality bar."""
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "torch-gpu", "unavailable": "no CUDA device"}), file=stdout, flush=True)
        return
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.deterministic = True
    frames = [synth_frame(i) for i in range(2)]
    ref = load_reference()
    mode = None
    if ref is not None:
        kind = "reference"
        td = tempfile.TemporaryDirectory()
        path = save_synth_model(td.name)
        try:
            torch.set_default_tensor_type(torch.cuda.HalfTensor)   # what run.py:382-383 does in fp16 mode
            mode = "set_default_tensor_type(cuda.HalfTensor), as run.py"
        except Exception as e:  # removed in some future torch: same numerics with explicit .half()
            mode = "explicit .half() (set_default_tensor_type failed: %s)" % type(e).__name__
        model = ref.Model(path, "infer", None, device=dev)
        model.model.half()
        from utils.utils import np2tensor, tensor2np

        def one(img):
            t = np2tensor(img).to(dev).half()
            return tensor2np(model(t).detach())
    else:
        from oracle import rrdb_oracle as O
        kind = "port"
        mode = "oracle functions on CUDA fp16 (baseline/_ref absent)"
        sd = {k: v.to(dev).half() for k, v in O.make_state_dict(scale=SCALE, nb=NB, nf=NF, seed=0).items()}

        def one(img):
            with torch.no_grad():
                return O.tensor2np(O.chop_forward(sd, O.np2tensor(img).to(dev).half(), patch_size=PATCH))
    steps = max(1, min(args.steps, 3))
    for _ in range(1):
        out = one(frames[0])
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    for i in range(steps):
        out = one(frames[i % 2])
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    out_pix = SCALE * H * SCALE * W
    from innfer_b200 import synth
    flop = 190 * PATCH * PATCH * synth.flop_per_lr_pixel(SCALE, NB, NF)
    line = {"impl": "torch-gpu", "metric": "output Mpix/s", "value": out_pix / sec / 1e6, "unit": "Mpix/s", "n_gpus": 1,
            "steps": steps, "warmup": 1, "ms_per_step": 1e3 * sec, "higher_is_better": True, "dtype": "f16",
            "data": "synthetic", "kind": kind, "mode": mode, "tflops": flop / sec / 1e12,
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1,
                       "timing": "host wall clock around whole frames (np2tensor, H2D, 190-tile loop, recompose, D2H, tensor2np)"},
            "clocks": clocks, "checksum": int(np.asarray(out)[::97, ::89].astype(np.int64).sum())}
    print(json.dumps(line), file=stdout, flush=True)


def torch_gpu_subprocess(args):
    """Runs `bench.py --impl torch-gpu` in a fresh process (the reference changes torch's default tensor type) and
    returns its JSON line as a dict, or {"unavailable": why}."""
    try:
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "torch-gpu", "--steps", "2"], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        for line in reversed(p.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": "no JSON line (rc=%d): %s" % (p.returncode, p.stderr.strip()[-300:])}
    except Exception as e:
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


# ------------------------------------------------------------------------------------------------ strong-scaling leg
def run_tile_sharded(args, eng, dist, rank, world, dev, warm, single_ms, ref_sha):
    """All ranks split the tiles of every frame (innfer_b200/multi_gpu.py).  Frame f is seed f % 2 of rank 0's
    single-GPU frames, so the stitched result can be compared by hash.  Returns the `strong` object on rank 0."""
    from innfer_b200 import multi_gpu as MG
    be = MG.NativeTileBackend(eng, H, W, PATCH, STEP)
    up = MG.TileShardedUpscaler(be, dist)
    pinned = [torch.from_numpy(synth_frame(i)).pin_memory() for i in range(2)]
    out_pix = SCALE * H * SCALE * W
    checks = []   # (frame, identical?)

    def frame_for(f):
        return pinned[f % 2] if MG.frame_owner(f, world) == rank else None

    def sha_of(f):
        res = up.result(f)
        if res is not None:
            checks.append((f, hashlib.sha256(res.tobytes()).hexdigest() == ref_sha[f % 2]))

    nwarm = max(warm, world)   # every rank owns (allocates, uploads, blends, downloads) at least once before timing
    for f in range(nwarm):
        up.submit(f, frame_for(f))
        sha_of(f)              # untimed: fetch and hash every warm-up frame
    torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index)
    sampler.start(settle=0.7)
    dist.barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s in (be.s_up, be.s_bl):
        s.wait_stream(be.s_cmp)
    e0.record(be.s_cmp)
    base = nwarm
    for i in range(args.steps):
        f = base + i
        up.submit(f, frame_for(f))
        g = f - world          # results are read one owned frame behind, so the device always has work queued
        if g >= base:
            up.result(g)
    last = list(range(max(base, base + args.steps - world), base + args.steps))
    for s in (be.s_up, be.s_bl):
        be.s_cmp.wait_stream(s)
    e1.record(be.s_cmp)
    torch.cuda.synchronize(dev)
    ms_rank = e0.elapsed_time(e1)
    clocks = sampler.stop()
    for g in last:             # untimed: the last frames still sit in the pinned result slots
        sha_of(g)
    t = torch.tensor([ms_rank], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    gathered = [None] * world
    dist.all_gather_object(gathered, {"ms": ms_rank / args.steps, "sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"],
                                      "checks": checks})
    up.close()
    if rank != 0:
        return None
    allchecks = [c for g in gathered for c in g["checks"]]
    return {"scaling": "strong", "ms_per_frame": ms, "value": out_pix / (ms * 1e-3) / 1e6, "unit": "Mpix/s",
            "steps": args.steps, "warmup": nwarm,
            "speedup_vs_1gpu": single_ms / ms, "efficiency_vs_1gpu": single_ms / (ms * world),
            "single_gpu_ms_per_frame": single_ms,
            "bit_identical": bool(allchecks) and all(ok for _, ok in allchecks), "frames_checked": len(allchecks),
            "per_rank_ms": [g["ms"] for g in gathered], "per_rank_sm_mhz": [g["sm_mhz"] for g in gathered],
            "per_rank_reasons": [g["reasons"] for g in gathered],
            "tiles_per_rank": [b - a for a, b in MG.partition(be.ntiles, world)],
            "h2d_bytes_per_frame": H * W * 3, "d2h_bytes_per_frame": out_pix * 3,
            "p2p_bytes_per_frame": int(be.ntiles * be.tile_bytes * (world - 1) / world) + (world - 1) * H * W * 3,
            "sharding": "tiles of one frame across ranks; last conv stores tiles into the owner's buffer through CUDA-IPC "
                        "peer mappings (NVLink); device-flag pipeline (upload / compute / blend streams), owner rotates, "
                        "blend + D2H of frame f overlap the compute of frame f+1; no NCCL on the data path",
            "timing": "CUDA events on every rank from after the start barrier to the completion of all three streams, max over "
                      "ranks; pinned-host H2D of every frame and D2H of every result inside the timed region"}


# ------------------------------------------------------------------------------------------------ our arm
def family_roofline(eng, frame_fn, peaks):
    """One extra frame with events around every conv launch (innfer_rrdb_profile mode 2): per kernel family
    launches, average ms, algorithmic FLOP and bytes per launch and the fractions of the two measured peaks."""
    eng.profile_reset(2)
    frame_fn()
    torch.cuda.synchronize()
    fam = eng.profile_families()
    eng.profile_reset(0)
    metrics = {}
    mpath = os.path.join(ROOT, "profiles", "r02_conv_metrics.json")
    if os.path.exists(mpath):
        with open(mpath) as f:
            metrics = json.load(f)
    total_ms = sum(v[1] for v in fam.values()) or 1.0
    out = []
    for name, (n, ms, flop, nbytes) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        e = {"kernel": name, "launches_per_step": n, "avg_ms": ms / n, "share_of_conv_time": ms / total_ms,
             "flop_per_launch": flop / n, "tflops": flop / (ms * 1e-3) / 1e12,
             "frac_of_sustained_tensor_peak": flop / (ms * 1e-3) / 1e12 / float(peaks["bf16_tflops_sustained"]),
             "bytes_per_launch": nbytes / n, "gbs": nbytes / (ms * 1e-3) / 1e9,
             "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"])}
        m = metrics.get("families", {}).get(name)
        if m:   # from the committed ncu capture of this same workload (profiles/), not measured live
            e["ncu"] = m
        out.append(e)
    return out


def run_chain(args, dev, warm):
    """BASELINE configs[2]: 1x RRDB + 4x RRDB + -cf on 1280x720 frames through innfer_b200.run.ChainRunner (the device
    pipeline the CLI uses)."""
    from innfer_b200 import _native as N
    from innfer_b200 import synth
    from innfer_b200.engine import RRDBEngine
    from innfer_b200.run import ChainRunner
    wl = WORKLOADS["chain"]
    h, w = wl["H"], wl["W"]
    engines = []
    for s in wl["chain"]:
        sd = synth.make_state_dict(scale=s, nb=NB, nf=NF, seed=s)
        engines.append(RRDBEngine.from_state_dict(sd, dict(in_nc=3, out_nc=3, nf=NF, nb=NB, gc=32, scale=s, plus=False), dev, fp16=True))
    runner = ChainRunner(engines, dev, cf=wl["cf"], patch_size=PATCH, step=STEP)
    frames = [synth_frame(10 + i, h, w) for i in range(2)]
    d_in = [torch.from_numpy(f).to(dev) for f in frames]
    out_pix = SCALE * h * SCALE * w
    for i in range(warm):
        runner.run_device(d_in[i % 2])
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    sampler.start()
    for e in engines:
        e.profile_reset(1)
    l0 = N.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = runner.run_device(d_in[i % 2])
    e1.record()
    torch.cuda.synchronize()
    launches = N.kernel_launches() - l0
    ms = e0.elapsed_time(e1) / args.steps
    conv_ms = sum(e.profile_read()[0] for e in engines)
    conv_launches = sum(e.profile_read()[1] for e in engines)
    for e in engines:
        e.profile_reset(0)
    clocks = sampler.stop()
    for i in range(2):
        res = runner(frames[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        res = runner(frames[i % 2])
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    peaks, peak_src = measured_peaks()
    flop = wl["tiles"] * PATCH * PATCH * sum(synth.flop_per_lr_pixel(s, NB, NF) for s in wl["chain"])
    achieved = flop * args.steps / (conv_ms * 1e-3) / 1e12
    peak = float(peaks["bf16_tflops_sustained"])
    line = {"metric": "output Mpix/s", "value": out_pix / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": 1,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": wl["name"], "frames_per_step_per_gpu": 1,
                       "l2": "per-step working set (GBs of tile activations) >> 126 MB L2, no flush needed",
                       "checksum": int(np.asarray(res)[::97, ::89].astype(np.int64).sum())},
            "e2e": {"value": out_pix / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h * w * 3, "d2h_bytes_per_step": out_pix * 3,
                    "timing": "host wall clock around ChainRunner.__call__ (numpy frame in -> numpy frame out, the call run.py makes)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "all %d conv launches of a step (both networks)" % (conv_launches // args.steps),
                         "flop_per_launch_avg": flop * args.steps / conv_launches, "avg_launch_ms": conv_ms / conv_launches,
                         "peak_source": peak_src},
            "cpu_baseline": None}
    print(json.dumps(line), flush=True)


def run_small(args, dev, warm):
    """BASELINE configs[3]: the small SR families on 512x512 frames through run.Model / ChainRunner (the CLI's device loop);
    beside each the unmodified reference (baseline/_ref) on the same GPU: its Model.__call__ (per-tile loop, PyTorch-dispatched
    cuDNN fp16) wrapped in np2tensor / tensor2np as its main() does."""
    from innfer_b200 import _native as N
    from innfer_b200 import run as R
    from innfer_b200.architectures import get_network
    from innfer_b200.utils.defaults import get_network_G_config
    torch.backends.cudnn.benchmark = True
    ref = load_reference()
    size = 512
    frames = [synth_frame(20 + i, size, size) for i in range(2)]
    d_in = [torch.from_numpy(f).to(dev) for f in frames]
    out_pix = SCALE * size * SCALE * size
    cases = []
    sampler = ClockSampler(dev.index)
    sampler.start()
    launches0 = N.kernel_launches()
    td = tempfile.TemporaryDirectory()
    for arch in ("srgan", "pan", "ppon"):
        torch.manual_seed(0)
        net = get_network(get_network_G_config({"type": arch}, SCALE))
        path = os.path.join(td.name, "%dx_rand_%s.pth" % (SCALE, arch))
        torch.save(net.state_dict(), path)
        m = R.Model(path, "infer", SCALE if arch != "srgan" else None, device=dev)
        m.model.half()
        runner = R.ChainRunner(R.native_chain([m], dev, True), dev)
        for i in range(max(warm, 3)):
            runner.run_device(d_in[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            runner.run_device(d_in[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        for i in range(2):
            res = runner(frames[i % 2])
        t0 = time.perf_counter()
        for i in range(args.steps):
            res = runner(frames[i % 2])
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        rec = {"net": arch, "ms": ms, "mpix_s": out_pix / ms / 1e3, "e2e_ms": e2e_ms,
               "checksum": int(np.asarray(res)[::97, ::89].astype(np.int64).sum())}
        if ref is not None and not args.no_torch_gpu:
            from utils.utils import np2tensor, tensor2np      # the reference's own (baseline/_ref on sys.path)
            rm = ref.Model(path, "infer", SCALE if arch != "srgan" else None, device=dev)
            rm.model.half()

            def one(img):
                return tensor2np(rm(np2tensor(img).to(dev).half()).detach())
            for i in range(2):            # cudnn.benchmark autotunes per shape: keep that out of the timed frames
                one(frames[i % 2])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(3):
                out = one(frames[i % 2])
            torch.cuda.synchronize()
            ref_ms = (time.perf_counter() - t0) * 1e3 / 3
            rec["torch_gpu_reference_ms"] = ref_ms
            rec["ours_e2e_over_reference_gpu"] = ref_ms / e2e_ms
            # same frame through both (the timed loop above ends on frames[0])
            rec["max_abs_u8_diff_vs_reference_gpu_fp16"] = int(np.abs(np.asarray(one(frames[1])).astype(int) - np.asarray(runner(frames[1])).astype(int)).max())
            del rm
        cases.append(rec)
        del runner, m
    clocks = sampler.stop()
    head = cases[0]
    line = {"metric": "output Mpix/s", "value": head["mpix_s"], "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(warm, 3), "ms_per_step": head["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOADS["small"]["name"],
                       "l2": "25 tiles of activations (hundreds of MB for the 64-channel nets) exceed L2; two alternating frames"},
            "e2e": {"value": out_pix / head["e2e_ms"] / 1e3, "unit": "Mpix/s", "ms_per_step": head["e2e_ms"],
                    "h2d_bytes_per_step": size * size * 3, "d2h_bytes_per_step": out_pix * 3},
            "gpu_launches": int(N.kernel_launches() - launches0), "clocks": clocks, "roofline": None, "cases": cases,
            "cpu_baseline": None}
    print(json.dumps(line), flush=True)


def run_i2i(args, dev, warm):
    """BASELINE configs[4]: the two image-to-image generators at 256x256 and 1024x1024 through the nn.Module mirrors (the
    objects run.Model holds); beside each the same module's torch ops on this GPU (PyTorch-dispatched cuDNN, fp16)."""
    from innfer_b200 import _native as N
    from innfer_b200 import synth
    from innfer_b200.architectures import get_network
    from innfer_b200.utils.defaults import get_network_G_config
    torch.backends.cudnn.benchmark = True
    peaks, peak_src = measured_peaks()
    peak = float(peaks["bf16_tflops_sustained"])
    cases, head = [], None
    sampler = ClockSampler(dev.index)
    sampler.start()
    launches0 = N.kernel_launches()
    for family, arch in (("unet", "unet_256"), ("resnet", "resnet_9blocks")):
        torch.manual_seed(0)
        net = get_network(get_network_G_config({"type": arch}, 1)).train(family == "unet").to(dev).half()
        for size in (256, 1024):
            x_host = (torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(size)) * 2 - 1).half().pin_memory()
            y_host = torch.empty(1, 3, size, size, dtype=torch.float16).pin_memory()
            x = x_host.to(dev)
            flop = synth.i2i_flop(family, size, size)
            rec = {"net": arch, "size": size, "flop": flop}
            with torch.no_grad():
                for name, fn in (("ours", lambda t: net(t)), ("torch_gpu", lambda t: net.model(t.clone()))):
                    for _ in range(max(warm, 3)):
                        fn(x)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.steps):
                        y = fn(x)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.steps
                    rec[name] = {"ms": ms, "mpix_s": size * size / ms / 1e3, "tflops": flop / ms / 1e9}
                    if name == "ours":
                        rec["frac_of_sustained_tensor_peak"] = flop / ms / 1e9 / peak
                        t0 = time.perf_counter()
                        for _ in range(args.steps):       # end to end: pinned host tensor in, pinned host tensor out
                            y_host.copy_(net(x_host.to(dev, non_blocking=True)), non_blocking=True)
                            torch.cuda.current_stream().synchronize()
                        rec["e2e_ms"] = (time.perf_counter() - t0) * 1e3 / args.steps
                rec["ours_over_torch_gpu"] = rec["torch_gpu"]["ms"] / rec["ours"]["ms"]
                if family == "resnet":
                    # what `run.py -a resnet_9blocks` does per image: uint8 frame in, chop_forward over 200 px tiles
                    # (step 0.5), uint8 frame out -- ChainRunner with the [-1, 1] mapping folded into the engine
                    from innfer_b200.run import ChainRunner
                    frame = synth_frame(size, size, size)
                    runner = ChainRunner([net.native_engine(dev, torch.float16, unit_io=True)], dev)
                    for _ in range(2):
                        runner(frame)
                    t0 = time.perf_counter()
                    for _ in range(args.steps):
                        runner(frame)
                    rec["cli_chop_e2e_ms"] = (time.perf_counter() - t0) * 1e3 / args.steps
            cases.append(rec)
            if family == "resnet" and size == 1024:
                head = rec
        net.invalidate_engine()
        if not args.no_cpu_baseline:
            # the reference's -cpu mode for the same module (torch fp32 ops on the host cores), bounded: one 256x256 image
            cpu_net = net.float().cpu()
            xc = (torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(256)) * 2 - 1)
            with torch.no_grad():
                cpu_net.model(xc.clone())
                t0 = time.perf_counter()
                for _ in range(3):
                    cpu_net.model(xc.clone())
                cpu_ms = (time.perf_counter() - t0) * 1e3 / 3
            for rec in cases:
                if rec["net"] == arch:
                    rec["cpu_256_ms"] = cpu_ms
                    rec["cpu_cores"] = host_threads()
    clocks = sampler.stop()
    line = {"metric": "output Mpix/s", "value": head["ours"]["mpix_s"], "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(warm, 3), "ms_per_step": head["ours"]["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOADS["i2i"]["name"],
                       "l2": "activations of one 1024x1024 image (up to 134 MB per tensor) exceed L2 at the outer layers; "
                             "256x256 images are L2 resident (launch-latency bound)"},
            "e2e": {"value": 1024 * 1024 / head["e2e_ms"] / 1e3, "unit": "Mpix/s", "ms_per_step": head["e2e_ms"],
                    "h2d_bytes_per_step": 3 * 1024 * 1024 * 2, "d2h_bytes_per_step": 3 * 1024 * 1024 * 2},
            "gpu_launches": int(N.kernel_launches() - launches0), "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": head["ours"]["tflops"], "peak": peak, "unit": "TFLOP/s",
                         "frac": head["ours"]["tflops"] / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "all kernels of one resnet_9blocks forward at 1024x1024 (convolutions + normalisation)"},
            "cases": cases,
            "cpu_baseline": (None if args.no_cpu_baseline else
                             {"value": 256 * 256 / head["cpu_256_ms"] / 1e3, "unit": "Mpix/s", "cores": head["cpu_cores"],
                              "kind": "port", "sample": "resnet_9blocks, one 256x256 image through the nn.Module mirror's torch "
                              "fp32 modules on the host (the -cpu mode; identical module tree and arithmetic to the reference's)"})}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--workload", default="frame", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-tiles", type=int, default=3, help="tiles per step for the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu", action="store_true")
    ap.add_argument("--max-batch", type=int, default=0)
    ap.add_argument("--shard", default="both", choices=["both", "images", "tiles"],
                    help="N>1: 'images' = weak-scaling leg only, 'tiles' = strong-scaling leg only (needs a single-GPU "
                         "frame time: a short one is measured), 'both' (default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "torch-gpu":
        if rank == 0:
            real = sys.stdout
            with contextlib.redirect_stdout(sys.stderr):   # the reference prints; stdout carries the one JSON line
                run_torch_gpu(args, real)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU leg)")
    warm = max(3, args.warmup)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.workload == "chain":
        if world > 1:
            raise SystemExit("--workload chain is a single-GPU workload")
        run_chain(args, dev, warm)
        return
    if args.workload == "small":
        if world > 1:
            raise SystemExit("--workload small is a single-GPU workload")
        run_small(args, dev, warm)
        return
    if args.workload == "i2i":
        if world > 1:
            raise SystemExit("--workload i2i is a single-GPU workload")
        run_i2i(args, dev, warm)
        return
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

    def gather(obj):
        if dist is None:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # ---------------- device-resident leg ("value")
    steps_w = args.steps if args.shard != "tiles" else min(args.steps, 3)
    for i in range(warm):
        eng.upscale_u8_device(d_in[i % 2], PATCH, STEP, out=d_out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    eng.profile_reset(1)
    launches0 = N.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps_w):
        eng.upscale_u8_device(d_in[i % 2], PATCH, STEP, out=d_out)
    e1.record()
    barrier()
    launches = N.kernel_launches() - launches0
    my_ms = e0.elapsed_time(e1) / steps_w
    ms_per_step = max_over_ranks(my_ms)
    conv_ms, conv_launches = eng.profile_read()
    eng.profile_reset(0)
    clocks = sampler.stop()
    weak_ranks = gather({"ms": my_ms, "sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"]})
    value = world * out_pix / (ms_per_step * 1e-3) / 1e6

    # ---------------- end-to-end leg: pinned host uint8 in, pinned host uint8 out, copies timed
    for i in range(2):
        eng.upscale_u8(h_in[i % 2], PATCH, STEP, out=h_out)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps_w):
        eng.upscale_u8(h_in[i % 2], PATCH, STEP, out=h_out)
    t1.record()
    barrier()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1)) / steps_w
    e2e_value = world * out_pix / (e2e_ms * 1e-3) / 1e6
    checksum = int(h_out[::97, ::89].to(torch.int64).sum().item())

    # ---------------- strong-scaling leg (N > 1): tiles of every frame across all ranks
    strong = None
    if world > 1 and args.shard != "images":
        ref_sha = [None, None]
        if rank == 0:   # rank 0's frames are seeds 0 and 1: the single-GPU results the stitched frames must equal
            for i in range(2):
                eng.upscale_u8(h_in[i], PATCH, STEP, out=h_out)
                ref_sha[i] = hashlib.sha256(h_out.numpy().tobytes()).hexdigest()
        dist.broadcast_object_list(ref_sha, src=0)
        # e2e_ms: this run's own single-GPU frame time (pinned host in -> host out, max over the ranks working alone)
        strong = run_tile_sharded(args, eng, dist, rank, world, dev, warm, e2e_ms, ref_sha)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    flop_step = 190 * PATCH * PATCH * synth.flop_per_lr_pixel(SCALE, NB, NF)
    roof = None
    if conv_launches:
        achieved = flop_step * steps_w / (conv_ms * 1e-3) / 1e12
        peak = float(peaks["bf16_tflops_sustained"])
        # the same conv sequence seen from the HBM side: conv1..conv4 of every dense block (row-streaming
        # kernel, 190-240 FLOP per byte) run at the DRAM roofline when timed alone (profiles/)
        bytes_step = 190 * PATCH * PATCH * synth.bytes_per_lr_pixel(SCALE, NB, NF)
        hbm_peak = float(peaks["hbm_gbs"])
        hbm_achieved = bytes_step * steps_w / (conv_ms * 1e-3) / 1e9
        fams = family_roofline(eng, lambda: eng.upscale_u8_device(d_in[0], PATCH, STEP, out=d_out), peaks)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_conv_metrics.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dominant", {}).get("dram_bytes_per_launch")
        dom = fams[0] if fams else None
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": "all %d conv launches of a step (conv_rows / conv_rows_pair / conv_up / conv_tc), timed in the real "
                          "schedule (programmatic dependent launch on); `families` lists each kernel family timed launch by "
                          "launch in one extra frame" % (conv_launches // steps_w),
                "dominant_kernel": dom["kernel"] if dom else None,
                "flop_per_launch_avg": flop_step * steps_w / conv_launches,
                "avg_launch_ms": conv_ms / conv_launches, "peak_source": peak_src,
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "algorithmic_bytes_per_step": bytes_step},
                "families": fams}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        c = CpuReference(args.ref_tiles)
        c.step()                       # warm-up pass
        secs = [c.step() for _ in range(2)]
        cpu = {"value": c.rate(float(np.mean(secs))), "unit": "Mpix/s", "cores": c.threads, "kind": c.kind,
               "sample": c.sample() + " (1 warm-up + 2 timed passes, %.1f s)" % sum(secs)}
    tgpu = None
    if not args.no_torch_gpu and world == 1:
        t = torch_gpu_subprocess(args)
        tgpu = t if "unavailable" in t else {
            "value": t["value"], "unit": "Mpix/s", "ms_per_step": t["ms_per_step"], "tflops": t["tflops"], "kind": t["kind"],
            "mode": t["mode"], "clocks": t["clocks"], "timing": t["config"]["timing"],
            "ours_e2e_over_this": e2e_value / t["value"],
            "what": "the unmodified reference (PyTorch-dispatched cuDNN fp16, cudnn.benchmark, per-tile loop of "
                    "run.py:187-197) on this same GPU and frame, end to end"}
    line = {
        "metric": "output Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": steps_w,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "sharding": "by image, no collective",
                   "l2": "per-step working set (GBs of tile activations) >> 126 MB L2, no flush needed",
                   "checksum": checksum},
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": H * W * 3,
                "d2h_bytes_per_step": out_pix * 3},
        "gpu_launches": int(launches), "clocks": clocks,
        "per_rank_ms": [r["ms"] for r in weak_ranks], "per_rank_sm_mhz": [r["sm_mhz"] for r in weak_ranks],
        "roofline": roof, "cpu_baseline": cpu, "torch_gpu_baseline": tgpu, "strong": strong,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
