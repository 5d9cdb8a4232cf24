"""Bring-up / diagnostics script for the GPU box (not collected by pytest).

    python tests/gpu_bringup.py --stage conv1|convs|net|time

Each stage prints PASS/FAIL lines with error statistics so that one gpurun call tells which layer of
the stack (TMA box, UMMA descriptors, epilogue, schedule) is wrong.
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from innfer_b200 import _native as N  # noqa: E402
from oracle import rrdb_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def conv_case(cin, cout, h, w, n=1, up=1, lrelu=False, res=False, fp32=False, seed=0, verbose=True, wide=False):
    lib = N.load()
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, cin, h, w, generator=g) * 2 - 1)
    wgt = (torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) / np.sqrt(cin * 9.0) * 2
    b = torch.rand(cout, generator=g) - 0.5
    r = (torch.rand(n, cout, h * up, w * up, generator=g) * 2 - 1) if res else None
    dt = torch.float32 if fp32 else torch.float16
    xd = x.to(dev, dt)
    rd = r.to(dev, dt) if res else None
    y = torch.empty(n, cout, h * up, w * up, device=dev, dtype=dt)
    wc = wgt.contiguous().numpy()
    bc = b.contiguous().numpy()
    rc = lib.innfer_conv3x3(xd.data_ptr(), n, cin, h, w, wc.ctypes.data, bc.ctypes.data, cout, up,
                            int(lrelu), rd.data_ptr() if res else None, 0.2, y.data_ptr(),
                            N.INNFER_F32 if fp32 else N.INNFER_F16, 2 if wide else int(fp32), None)
    if rc != 0:
        print("FAIL rc=%d %s" % (rc, N.last_error()))
        return False
    torch.cuda.synchronize()
    # reference in fp64 on the (rounded) inputs the kernel saw
    xr = xd.double()
    wr = (wgt.to(dev, dt)).double() if not fp32 else wgt.to(dev).double()
    if up > 1:
        xr = F.interpolate(xr, scale_factor=float(up), mode="nearest")
    ref = F.conv2d(xr, wr, b.to(dev).double(), padding=1)
    if lrelu:
        ref = F.leaky_relu(ref, 0.2)
    if res:
        ref = ref * 0.2 + rd.double()
    err = (y.double() - ref).abs()
    tol = 2e-5 if fp32 else 2e-2
    scale = ref.abs().max().item()
    ok = bool(err.max().item() <= tol * max(scale, 1.0)) and bool(torch.isfinite(y).all())
    tag = "PASS" if ok else "FAIL"
    print("%s conv cin=%d cout=%d %dx%d n=%d up=%d lrelu=%d res=%d fp32=%d wide=%d: max_err=%.3e mean_err=%.3e ref_max=%.3f"
          % (tag, cin, cout, h, w, n, up, lrelu, res, fp32, wide, err.max().item(), err.mean().item(), scale))
    if not ok and verbose:
        bad = err > tol * max(scale, 1.0)
        print("   bad fraction %.4f" % bad.float().mean().item())
        e = err[0]
        print("   per-out-channel max err:", [round(v, 3) for v in e.amax(dim=(1, 2)).tolist()][:16])
        ey = e.amax(dim=(0, 2))
        ex = e.amax(dim=(0, 1))
        print("   per-row max err:", [round(v, 3) for v in ey.tolist()][:40])
        print("   per-col max err:", [round(v, 3) for v in ex.tolist()][:48])
        print("   y[0,0,:4,:8]  ", y[0, 0, :4, :8].float().cpu().numpy().round(3).tolist())
        print("   ref[0,0,:4,:8]", ref[0, 0, :4, :8].float().cpu().numpy().round(3).tolist())
    return ok


def stage_conv1():
    ok = conv_case(64, 32, 16, 40)
    ok &= conv_case(16, 32, 16, 8)
    return ok


def stage_convs():
    ok = True
    for cin, cout in ((64, 32), (96, 32), (128, 32), (160, 32), (192, 64)):
        ok &= conv_case(cin, cout, 40, 48, lrelu=cout == 32, res=cout == 64)
    ok &= conv_case(3, 64, 33, 47)
    ok &= conv_case(64, 3, 50, 70)
    ok &= conv_case(64, 64, 37, 53, n=3, lrelu=True)
    ok &= conv_case(64, 64, 24, 40, up=2, lrelu=True)
    ok &= conv_case(64, 64, 19, 21, up=2, lrelu=True)
    ok &= conv_case(64, 64, 17, 23, up=3, lrelu=True)
    ok &= conv_case(32, 32, 20, 20, lrelu=True)
    ok &= conv_case(64, 32, 200, 200, n=2, lrelu=True)
    # wide layout (production layout of the fp16 path)
    for cin, cout in ((64, 32), (96, 32), (128, 32), (160, 32), (192, 64)):
        ok &= conv_case(cin, cout, 40, 48, n=2, lrelu=cout == 32, res=cout == 64, wide=True)
    ok &= conv_case(96, 32, 33, 47, n=3, res=True, wide=True)
    ok &= conv_case(64, 32, 2, 127, n=5, wide=True)
    ok &= conv_case(64, 64, 24, 40, n=2, up=2, lrelu=True, wide=True)
    ok &= conv_case(64, 3, 50, 70, n=2, wide=True)
    # fp32 direct kernel
    ok &= conv_case(64, 32, 40, 48, lrelu=True, fp32=True)
    ok &= conv_case(192, 64, 21, 35, res=True, fp32=True)
    ok &= conv_case(3, 64, 33, 47, fp32=True)
    ok &= conv_case(64, 3, 30, 30, fp32=True)
    ok &= conv_case(64, 64, 19, 21, up=2, lrelu=True, fp32=True)
    ok &= conv_case(64, 64, 17, 23, up=3, lrelu=True, fp32=True)
    return ok


def make_handle(sd, fp16=True, scale=None):
    lib = N.load()
    p = O.infer_params(sd)
    cfg = N.RRDBCfg(p["in_nc"], p["out_nc"], p["nf"], p["nb"], 32, scale or p["scale"], 0, int(fp16))
    h = ctypes.c_void_p()
    N.check(lib.innfer_rrdb_create(ctypes.byref(cfg), 0, ctypes.byref(h)))
    for k, v in sd.items():
        a = v.detach().cpu().float().contiguous().numpy()
        shp = (ctypes.c_int64 * a.ndim)(*a.shape)
        N.check(lib.innfer_rrdb_load(h, k.encode(), a.ctypes.data, shp, a.ndim))
    N.check(lib.innfer_rrdb_finalize(h))
    return h


def stage_net():
    lib = N.load()
    ok = True
    for scale, nb, hw, fp16 in ((4, 2, (40, 56), True), (1, 2, (48, 40), True), (2, 1, (32, 32), True),
                                (4, 23, (64, 64), True), (4, 2, (40, 56), False), (3, 1, (24, 24), True)):
        sd = O.make_state_dict(scale=scale, nb=nb, seed=1)
        sdd = {k: v.to(dev) for k, v in sd.items()}
        h = make_handle(sd, fp16=fp16, scale=scale)
        img = np.random.default_rng(3).integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
        x = O.np2tensor(img).to(dev)
        ref = O.rrdbnet_forward(sdd, x, scale)
        dt = torch.float16 if fp16 else torch.float32
        xin = x.to(dt).contiguous()
        y = torch.empty(1, 3, scale * hw[0], scale * hw[1], device=dev, dtype=dt)
        rc = lib.innfer_rrdb_forward(h, xin.data_ptr(), 1, hw[0], hw[1], y.data_ptr(),
                                     N.INNFER_F16 if fp16 else N.INNFER_F32, None)
        if rc:
            print("FAIL forward rc=%d %s" % (rc, N.last_error()))
            ok = False
            continue
        torch.cuda.synchronize()
        err = (y.float() - ref).abs()
        u8a = O.tensor2np(y)
        u8b = O.tensor2np(ref)
        du8 = np.abs(u8a.astype(int) - u8b.astype(int))
        mse = np.mean((u8a.astype(float) - u8b.astype(float)) ** 2)
        psnr = 99.0 if mse == 0 else 10 * np.log10(255 ** 2 / mse)
        rel = (err.max() / ref.abs().max()).item()
        good = du8.max() <= 1 and psnr >= 50 and (fp16 or rel < 1e-4)
        ok &= bool(good)
        print("%s net scale=%d nb=%d %s fp16=%d: max_abs=%.3e rel=%.3e u8_maxdiff=%d psnr=%.1f unclipped=%.2f"
              % ("PASS" if good else "FAIL", scale, nb, hw, fp16, err.max().item(), rel, du8.max(), psnr,
                 ((u8b > 0) & (u8b < 255)).mean()))
        # chop path on the same image with a small patch
        y2 = torch.empty_like(y)
        rc = lib.innfer_rrdb_chop_forward(h, xin.data_ptr(), hw[0], hw[1], 32 if scale != 3 else 200, 0.5, y2.data_ptr(),
                                          N.INNFER_F16 if fp16 else N.INNFER_F32, None)
        if rc:
            print("FAIL chop rc=%d %s" % (rc, N.last_error()))
            ok = False
        else:
            torch.cuda.synchronize()
            sdc = {k: v for k, v in sd.items()}
            refc = O.chop_forward(sdc, x.cpu(), patch_size=32 if scale != 3 else 200,
                                  forward=lambda t: O.rrdbnet_forward(sdc, t, scale)) if scale != 3 else \
                O.rrdbnet_forward(sdc, x.cpu(), scale)
            e2 = (y2.float().cpu() - refc).abs()
            d2 = np.abs(O.tensor2np(y2).astype(int) - O.tensor2np(refc).astype(int))
            good = d2.max() <= 1
            ok &= bool(good)
            print("%s chop scale=%d nb=%d %s fp16=%d: max_abs=%.3e u8_maxdiff=%d"
                  % ("PASS" if good else "FAIL", scale, nb, hw, fp16, e2.max().item(), d2.max()))
        lib.innfer_rrdb_destroy(h)
    return ok


def stage_prof():
    """One 800x1000 frame (63 tiles) for ncu: launch list and full captures."""
    lib = N.load()
    sd = O.make_state_dict(scale=4, nb=23, seed=0)
    h = make_handle(sd, fp16=True)
    H, W = 800, 1000
    if os.environ.get("INNFER_MB_PROF"):
        lib.innfer_rrdb_set_max_batch(h, int(os.environ["INNFER_MB_PROF"]))
    img = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
    din = torch.from_numpy(img).to(dev)
    dout = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    N.check(lib.innfer_rrdb_upscale_u8_device(h, din.data_ptr(), H, W, 200, 0.5, dout.data_ptr(), None))
    torch.cuda.synchronize()
    return True


def stage_trace():
    """clock64 trace of CTA 0 of the rows kernel with INNFER_TRACE_NCH input chunks (default conv1 = 8)."""
    os.environ.setdefault("INNFER_TRACE_NCH", "8")
    lib = N.load()
    buf = torch.zeros(3072 + 148 * 8, dtype=torch.int64, device=dev)
    lib.innfer_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    stage_prof()
    lib.innfer_debug_set_trace(None)
    t = buf.cpu().numpy()
    c = t[3072:].reshape(148, 8)
    if c[:, 1].any():
        g0 = c[:, 0].min()
        print("per-CTA (cycles): prologue = kernel start -> issue loop, loop, tail = loop end -> exit; start skew from globaltimer (ns)")
        print("  prologue median %.0f max %.0f | loop median %.0f min %.0f max %.0f | tail median %.0f max %.0f | skew max %d ns | stages min %d max %d"
              % (np.median(c[:, 2] - c[:, 1]), (c[:, 2] - c[:, 1]).max(), np.median(c[:, 3] - c[:, 2]), (c[:, 3] - c[:, 2]).min(),
                 (c[:, 3] - c[:, 2]).max(), np.median(c[:, 4] - c[:, 3]), (c[:, 4] - c[:, 3]).max(), (c[:, 0] - g0).max(),
                 c[:, 5].min(), c[:, 5].max()))
        lead = c[c[:, 5] > 0]
        lt = np.sort(lead[:, 3] - lead[:, 2])
        print("  issuing CTAs %d: loop cycles min %.0f p10 %.0f median %.0f p90 %.0f max %.0f" % (
            len(lead), lt[0], lt[len(lt) // 10], lt[len(lt) // 2], lt[(9 * len(lt)) // 10], lt[-1]))
        per = (c[:, 3] - c[:, 2]) / np.maximum(c[:, 5], 1)
        print("  cycles per stage by CTA: min %.0f median %.0f max %.0f; total kernel cycles (max over CTAs) %.0f"
              % (per.min(), np.median(per), per.max(), (c[:, 4] - c[:, 1]).max()))
    m = t[:1024].reshape(256, 4)
    prod = t[1024:1280]
    epi = t[2048:2304]
    t0 = m[0, 0]
    print("stage: start  +firsthalf  +wait  +rest | dt_start | TMA issue | epi tfull   (cycles, relative)")
    for i in list(range(0, 24)) + list(range(100, 124)):
        if m[i, 0] == 0:
            break
        print("%3d: %8d %5d %5d %5d | %5d | %8d | %8d" % (i, m[i, 0] - t0, m[i, 1] - m[i, 0], m[i, 2] - m[i, 1], m[i, 3] - m[i, 2],
                                                  m[i, 0] - m[i - 1, 0] if i else 0, prod[i] - t0, epi[i] - t0))
    ep = np.stack([t[2048:2304], t[2304:2560], t[2560:2816]], 1)
    k = int((ep[:, 0] != 0).sum())
    if k > 40:
        d = ep[8:k - 2]
        print("epilogue warp 2 (cycles, median): tfull->released %.0f, store_row %.0f, row period %.0f" % (
            np.median(d[:, 1] - d[:, 0]), np.median(d[:, 2] - d[:, 1]), np.median(np.diff(d[:, 0]))))
        print("  store_row samples:", [int(v) for v in (d[:24, 2] - d[:24, 1])])
        print("  idle before tfull (prev store end -> tfull seen):", [int(v) for v in (d[1:25, 0] - d[:24, 2])])
    n = int((m[:, 0] != 0).sum())
    w = t[2560:3072].reshape(256, 2)
    print("ready counter seen at stages 100..123:", [int(v) for v in w[100:124, 0]])
    print("wait block split (median): full-wait %.0f, slot-wait %.0f, fence %.0f" % (
        np.median(w[1:n - 1, 0] - m[1:n - 1, 1]), np.median(w[1:n - 1, 1] - w[1:n - 1, 0]), np.median(m[1:n - 1, 2] - w[1:n - 1, 1])))
    d = np.diff(m[:n, 0])
    print("stages traced %d; cycles per stage: median %.0f mean %.0f; first-half issue median %.0f, wait median %.0f, rest median %.0f"
          % (n, np.median(d), d.mean(), np.median(m[:n, 1] - m[:n, 0]), np.median(m[:n, 2] - m[:n, 1]), np.median(m[:n, 3] - m[:n, 2])))
    return True


def stage_trace_pair():
    """clock64 trace of the conv5 CTA-pair epilogue (INNFER_TRACE_BUILD=1): warp 2 of CTA 0, per row
    tfull seen -> first block read -> store_row done -> slot released."""
    os.environ["INNFER_TRACE_NCH"] = "24"
    lib = N.load()
    buf = torch.zeros(3072 + 148 * 8, dtype=torch.int64, device=dev)
    lib.innfer_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    stage_prof()
    lib.innfer_debug_set_trace(None)
    t = buf.cpu().numpy()
    ep = np.stack([t[2048:2304], t[2304:2560], t[2560:2816], t[2816:3072]], 1)
    k = int((ep[:, 3] != 0).sum())
    d = ep[8:k - 2]
    print("rows traced", k)
    print("median cycles: row period %.0f | tfull->block2 read %.0f | store_row %.0f | blocks 1,0 + release %.0f | released->next tfull %.0f"
          % (np.median(np.diff(d[:, 0])), np.median(d[:, 1] - d[:, 0]), np.median(d[:, 2] - d[:, 1]),
             np.median(d[:, 3] - d[:, 2]), np.median(d[1:, 0] - d[:-1, 3])))
    for i in range(8, 40):
        print("%3d: period %6d | ld2 %5d | store %5d | ld1,0+rel %5d | idle %6d" % (
            i, ep[i + 1, 0] - ep[i, 0], ep[i, 1] - ep[i, 0], ep[i, 2] - ep[i, 1], ep[i, 3] - ep[i, 2], ep[i + 1, 0] - ep[i, 3]))
    c = t[3072:].reshape(148, 8)
    lead = c[c[:, 5] > 0]
    if len(lead):
        per = (lead[:, 3] - lead[:, 2]) / lead[:, 5]
        print("issuing CTAs %d: cycles per stage (18 MMAs) median %.0f min %.0f max %.0f" % (len(lead), np.median(per), per.min(), per.max()))
    return True


def stage_steady():
    """Power-limited steady state of single conv kernels: ~3 s of back-to-back launches each, NVML sampled."""
    import threading
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    lib = N.load()
    B = int(os.environ.get("INNFER_STEADY_B", "95"))
    cases = ((64, 32, 0), (96, 32, 0), (128, 32, 0), (160, 32, 0), (192, 64, 1), (64, 64, 0))
    if os.environ.get("INNFER_STEADY_CASES"):
        cases = tuple(tuple(int(v) for v in c.split(":")) for c in os.environ["INNFER_STEADY_CASES"].split(","))
    hw = int(os.environ.get("INNFER_STEADY_HW", "200"))
    for cin, cout, res in cases:
        flop = 2.0 * 9 * cin * cout * B * hw * hw
        ms = ctypes.c_float(0)
        N.check(lib.innfer_debug_conv_loop(cin, cout, B, hw, hw, res, 20, 50, ctypes.byref(ms)))
        iters = max(50, int(float(os.environ.get('INNFER_STEADY_MS', '3000')) / (ms.value / 50)))
        samples, stop = [], threading.Event()

        def sampler():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                                pynvml.nvmlDeviceGetPowerUsage(hnd) / 1e3,
                                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)))
                time.sleep(0.05)
        th = threading.Thread(target=sampler)
        th.start()
        N.check(lib.innfer_debug_conv_loop(cin, cout, B, hw, hw, res, 20, iters, ctypes.byref(ms)))
        stop.set()
        th.join()
        tail = samples[len(samples) // 2:]   # second half of the run: thermally / power settled
        clk = sorted(v[0] for v in tail)[len(tail) // 2]
        pw = sorted(v[1] for v in tail)[len(tail) // 2]
        reasons = 0
        for v in tail:
            reasons |= v[2]
        us = ms.value / iters * 1e3
        print("steady conv %3d->%2d res=%d B=%d: %.1f us/launch  %.0f TFLOP/s  sm %d MHz  %.0f W  reasons 0x%x  tensor util at that clock %.0f%%"
              % (cin, cout, res, B, us, flop / us / 1e6, clk, pw, reasons, 100 * flop / us / 1e6 / (148 * 8192 * clk / 1e6)))
    return True


def stage_ppon():
    """PPON engine vs the oracle (small nets), fp16 and fp32, plus an intermediate check of one residual block."""
    from innfer_b200.engine import PPONEngine
    ok = True
    for scale, nb, hw, fp16 in ((4, 1, (40, 48), True), (2, 2, (36, 44), True), (4, 1, (40, 48), False), (1, 1, (33, 47), True),
                                (3, 1, (24, 24), True)):
        sd = O.make_ppon_state_dict(scale=scale, nb=nb, seed=17)
        eng = PPONEngine.from_state_dict(sd, dict(in_nc=3, out_nc=3, nf=64, nb=nb, scale=scale, alpha=1.0), dev, fp16=fp16)
        img = np.random.default_rng(18).integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
        x = O.np2tensor(img)
        ref = O.ppon_forward(sd, x, scale)[2]
        y = eng.forward(x.to(dev).half() if fp16 else x.to(dev)).float().cpu()
        err = (y - ref).abs()
        du8 = np.abs(O.tensor2np(y).astype(int) - O.tensor2np(ref).astype(int))
        rel = (err.max() / ref.abs().max()).item()
        good = du8.max() <= 1 and (fp16 or rel < 1e-4) and bool(torch.isfinite(y).all())
        ok &= bool(good)
        print("%s ppon scale=%d nb=%d %s fp16=%d: max_abs=%.3e rel=%.3e u8_maxdiff=%d" % (
            "PASS" if good else "FAIL", scale, nb, hw, fp16, err.max().item(), rel, du8.max()))
        yc = eng.chop_forward(x.to(dev).half() if fp16 else x.to(dev), 32, 0.5).float().cpu()
        refc = O.chop_forward(sd, x, patch_size=32, scale=scale, forward=lambda t: O.ppon_forward(sd, t, scale)[2])
        d2 = np.abs(O.tensor2np(yc).astype(int) - O.tensor2np(refc).astype(int))
        good = d2.max() <= 1
        ok &= bool(good)
        print("%s ppon chop scale=%d: u8_maxdiff=%d" % ("PASS" if good else "FAIL", scale, d2.max()))
        eng.close()
    return ok


def stage_ppon_time():
    """PPON at the reference's default depth (nb = 24) on a 1920x1080 frame: device time per frame."""
    from innfer_b200.engine import PPONEngine
    sd = O.make_ppon_state_dict(scale=4, nb=24, seed=0)
    eng = PPONEngine.from_state_dict(sd, dict(in_nc=3, out_nc=3, nf=64, nb=24, scale=4, alpha=1.0), dev, fp16=True)
    H, W = 1080, 1920
    din = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)).to(dev)
    dout = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    mac = 9 * 3 * 64 + 28 * 3 * (9 * 64 * 64 + 8 * 9 * 64 * 32 + 256 * 64) + 9 * 64 * 64 \
        + 3 * (4 * 9 * 64 * 64 + 16 * 9 * 64 * 64 + 16 * 9 * 64 * 64 + 16 * 9 * 64 * 3)
    flop = 190 * 200 * 200 * 2.0 * mac
    for mb in [int(v) for v in os.environ.get("INNFER_MB", "95").split(",")]:
        eng.set_max_batch(mb)
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.upscale_u8_device(din, 200, 0.5, out=dout)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("ppon 1080p nb=24 max_batch=%d iter=%d: %.1f ms  %.1f out-Mpix/s  %.1f TFLOP/s (%.1f TFLOP per frame)" %
                  (mb, it, ms, 16 * H * W / ms / 1e3, flop / ms / 1e9, flop / 1e12))
    print("out mean", dout.float().mean().item(), "launches", N.kernel_launches())
    eng.close()
    return True


def stage_pan_time():
    """PAN at the reference's defaults (nf 40, unf 24, nb 16, self attention) on a 1920x1080 frame, fp16 and fp32:
    device time per frame, and the share of the attention kernels (profile of one 95-tile batch via CUDA events)."""
    from innfer_b200.engine import PANEngine
    sd = O.make_pan_state_dict(scale=4, nb=16, seed=0)
    H, W = 1080, 1920
    din = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)).to(dev)
    dout = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    for fp16 in (True, False):
        eng = PANEngine.from_state_dict(sd, dict(in_nc=3, out_nc=3, nf=40, unf=24, nb=16, scale=4, self_attention=True,
                                                 double_scpa=False), dev, fp16=fp16)
        for it in range(3):
            l0 = N.kernel_launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.upscale_u8_device(din, 200, 0.5, out=dout)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("pan 1080p %s iter=%d: %.1f ms  %.1f out-Mpix/s  (%d launches)" %
                  ("fp16" if fp16 else "fp32", it, ms, 16 * H * W / ms / 1e3, N.kernel_launches() - l0))
        print("out mean", dout.float().mean().item())
        eng.close()
    return True


def stage_lat():
    """Latency of small single-tile inputs (BASELINE configs[0] size, 64x64, and one 200x200 tile): launch-gap bound,
    which is what programmatic dependent launch addresses (A/B with INNFER_PDL=0)."""
    from innfer_b200.engine import PANEngine
    sd = O.make_state_dict(scale=4, nb=23, seed=0)
    h = make_handle(sd, fp16=True)
    lib = N.load()
    psd = O.make_pan_state_dict(scale=4, nb=16, seed=0)
    pan = PANEngine.from_state_dict(psd, dict(in_nc=3, out_nc=3, nf=40, unf=24, nb=16, scale=4, self_attention=True,
                                              double_scpa=False), dev, fp16=True)
    for hw in (64, 200):
        din = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (hw, hw, 3), dtype=np.uint8)).to(dev)
        dout = torch.empty(4 * hw, 4 * hw, 3, dtype=torch.uint8, device=dev)
        for name in ("rrdb", "pan"):
            best, host = 1e9, 1e9
            for it in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                t0 = time.perf_counter()
                if name == "rrdb":
                    lib.innfer_rrdb_upscale_u8_device(h, din.data_ptr(), hw, hw, 200, 0.5, dout.data_ptr(), None)
                else:
                    pan.upscale_u8_device(din, 200, 0.5, out=dout)
                t1 = time.perf_counter()
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    best = min(best, e0.elapsed_time(e1))
                    host = min(host, (t1 - t0) * 1e3)
            print("lat %s %dx%d: %.3f ms on the device, %.3f ms of host enqueue" % (name, hw, hw, best, host))
    pan.close()
    return True


def stage_srres_time():
    """SRResNet at the reference's defaults (nf 64, nb 16, pixel-shuffle upsampler) on a 1920x1080 frame, fp16."""
    from innfer_b200.engine import SRResNetEngine
    sd = O.make_srresnet_state_dict(scale=4, nb=16, seed=0)
    eng = SRResNetEngine.from_state_dict(sd, dict(in_nc=3, out_nc=3, nf=64, nb=16, scale=4, upsample_mode="pixelshuffle",
                                                  res_scale=1.0), dev, fp16=True)
    H, W = 1080, 1920
    din = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)).to(dev)
    dout = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    mac = 9 * 3 * 64 + 33 * 9 * 64 * 64 + 9 * 64 * 256 + 4 * 9 * 64 * 256 + 16 * 9 * 64 * 64 + 16 * 9 * 64 * 3
    flop = 190 * 200 * 200 * 2.0 * mac
    for it in range(3):
        l0 = N.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.upscale_u8_device(din, 200, 0.5, out=dout)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("srresnet 1080p nb=16 iter=%d: %.1f ms  %.1f out-Mpix/s  %.1f TFLOP/s  (%d launches)" %
              (it, ms, 16 * H * W / ms / 1e3, flop / ms / 1e9, N.kernel_launches() - l0))
    eng.close()
    return True


def stage_trace_up():
    """clock64 trace of CTA 0 of the last conv_up launch of a frame (needs an INNFER_TRACE_BUILD=1 build)."""
    lib = N.load()
    buf = torch.zeros(3072 + 148 * 8, dtype=torch.int64, device=dev)
    lib.innfer_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    os.environ["INNFER_TRACE_NCH"] = "999"
    stage_prof()
    lib.innfer_debug_set_trace(None)
    t = buf.cpu().numpy()
    m = t[:1024].reshape(256, 4)
    e = t[2048:2048 + 128].reshape(64, 2)
    t0 = m[0, 0]
    print("stage: start | poll | issue | dt | (every 4th stage: epilogue tfull seen, set released)")
    for i in range(8, 48):
        extra = ""
        if i % 4 == 0:
            extra = " | tile %d epi tfull %d released %d" % (i // 4, e[i // 4, 0] - t0, e[i // 4, 1] - t0)
        print("%3d: %8d %5d %5d | %5d%s" % (i, m[i, 0] - t0, m[i, 1] - m[i, 0], m[i, 2] - m[i, 1], m[i, 0] - m[i - 1, 0], extra))
    return True


def stage_ppon_prof():
    """One 800x1000 frame (63 tiles) through a 2-block PPON for ncu launch lists."""
    from innfer_b200.engine import PPONEngine
    sd = O.make_ppon_state_dict(scale=4, nb=2, seed=0)
    eng = PPONEngine.from_state_dict(sd, dict(in_nc=3, out_nc=3, nf=64, nb=2, scale=4, alpha=1.0), dev, fp16=True)
    H, W = 800, 1000
    din = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)).to(dev)
    dout = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    eng.upscale_u8_device(din, 200, 0.5, out=dout)
    torch.cuda.synchronize()
    eng.close()
    return True


def stage_pan_prof():
    """One 800x1000 frame (63 tiles) through the default PAN (nb = 16) for ncu launch lists."""
    from innfer_b200.engine import PANEngine
    sd = O.make_pan_state_dict(scale=4, nb=16, seed=0)
    eng = PANEngine.from_state_dict(sd, dict(in_nc=3, out_nc=3, nf=40, unf=24, nb=16, scale=4, self_attention=True,
                                             double_scpa=False), dev, fp16=True)
    H, W = 800, 1000
    din = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)).to(dev)
    dout = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    eng.upscale_u8_device(din, 200, 0.5, out=dout)
    torch.cuda.synchronize()
    eng.close()
    return True


def stage_pix():
    """HBM-bound kernels at benchmark sizes: image->tiles, blend (+uint8), colour fix (for ncu)."""
    lib = N.load()
    H, W, s = 1080, 1920, 4
    img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)).to(dev)
    nt = 190
    tiles_in = torch.empty(nt, 2, 200, 200, 8, dtype=torch.float16, device=dev)
    tiles_out = torch.rand(nt, 1, 800, 800, 8, device=dev).half()
    out = torch.empty(s * H, s * W, 3, dtype=torch.uint8, device=dev)
    outf = torch.empty(1, 3, s * H, s * W, dtype=torch.float16, device=dev)
    # the engine's own path (compact 8-byte tile pixels): one whole frame of a 1-block net
    sd = O.make_state_dict(scale=4, nb=1, seed=0)
    h = make_handle(sd, fp16=True)
    N.check(lib.innfer_rrdb_upscale_u8_device(h, img.data_ptr(), H, W, 200, 0.5, out.data_ptr(), None))
    for _ in range(3):
        N.check(lib.innfer_image_to_tiles(img.data_ptr(), N.INNFER_U8, 3, H, W, 200, 0.5, tiles_in.data_ptr(), None))
        N.check(lib.innfer_blend(tiles_out.data_ptr(), H, W, 200, 0.5, s, 3, out.data_ptr(), N.INNFER_U8, None))
        N.check(lib.innfer_blend(tiles_out.data_ptr(), H, W, 200, 0.5, s, 3, outf.data_ptr(), N.INNFER_F16, None))
    h, w = 720, 1280
    lr = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (h, w, 3), dtype=np.uint8)).to(dev)
    sr = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (4 * h, 4 * w, 3), dtype=np.uint8)).to(dev)
    cf = torch.empty_like(sr)
    for _ in range(3):
        N.check(lib.innfer_color_fix(lr.data_ptr(), h, w, sr.data_ptr(), 4 * h, 4 * w, cf.data_ptr(), None))
    torch.cuda.synchronize()
    return True


def stage_cfg3():
    """BASELINE configs[2] at full size: chained 1x RRDB (JPEG-denoise stand-in) + 4x RRDB with -cf on 1280x720."""
    import time as _t
    from innfer_b200.engine import RRDBEngine
    from innfer_b200.utils import utils as U
    H, W = 720, 1280
    sd1 = O.make_state_dict(scale=1, nb=23, seed=5)
    sd4 = O.make_state_dict(scale=4, nb=23, seed=6)
    e1 = RRDBEngine.from_state_dict(sd1, dict(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, scale=1, plus=False), dev)
    e4 = RRDBEngine.from_state_dict(sd4, dict(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, scale=4, plus=False), dev)
    img = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
    lib = N.load()
    d_lr = torch.from_numpy(img).to(dev)
    d_cf = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = _t.perf_counter()
        x = O.np2tensor(img).to(dev).half()
        y = e4.chop_forward(e1.chop_forward(x, 200, 0.5), 200, 0.5)
        # tensor2np on the device: clip(255x).round() -> uint8 HWC BGR
        sr = (y[0].float().flip(0).permute(1, 2, 0) * 255).clamp_(0, 255).round_().to(torch.uint8).contiguous()
        N.check(lib.innfer_color_fix(d_lr.data_ptr(), H, W, sr.data_ptr(), 4 * H, 4 * W, d_cf.data_ptr(), None))
        out = d_cf.cpu()
        torch.cuda.synchronize()
        ms = (_t.perf_counter() - t0) * 1e3
        print("cfg3 720p 1x+4x+cf iter=%d: %.1f ms  %.1f out-Mpix/s (84+84 tiles, %.1f TFLOP)" %
              (it, ms, 16 * H * W / ms / 1e3, 84 * 40000 * (O.flop_per_lr_pixel(1) + O.flop_per_lr_pixel(4)) / 1e12))
    print("unclipped fraction", float(((out > 0) & (out < 255)).float().mean()))
    return True


def stage_time():
    lib = N.load()
    sd = O.make_state_dict(scale=4, nb=23, seed=0)
    h = make_handle(sd, fp16=True)
    H, W = 1080, 1920
    img = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
    din = torch.from_numpy(img).to(dev)
    dout = torch.empty(4 * H, 4 * W, 3, dtype=torch.uint8, device=dev)
    flop = 190 * 200 * 200 * O.flop_per_lr_pixel(4, 23, 64)
    for mb in [int(v) for v in os.environ.get('INNFER_MB', '38,19,10').split(',')]:
        lib.innfer_rrdb_set_max_batch(h, mb)
        for it in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.innfer_rrdb_upscale_u8_device(h, din.data_ptr(), H, W, 200, 0.5, dout.data_ptr(), None)
            e1.record()
            torch.cuda.synchronize()
            if rc:
                print("FAIL rc=%d %s" % (rc, N.last_error()))
                return False
            ms = e0.elapsed_time(e1)
            print("time 1080p max_batch=%d iter=%d: %.1f ms  %.1f out-Mpix/s  %.1f TFLOP/s" %
                  (mb, it, ms, 16 * H * W / ms / 1e3, flop / ms / 1e9))
    print("launches", N.kernel_launches())
    print("out mean", dout.float().mean().item())
    return True


def _i2i_net(family, ngf=64, seed=0):
    from innfer_b200.architectures import get_network
    from innfer_b200.utils.defaults import get_network_G_config
    torch.manual_seed(seed)
    net = get_network(get_network_G_config({"type": "unet_256" if family == "unet" else "resnet_9blocks", "ngf": ngf}, 1))
    return net.train(family == "unet")     # run.py:295-309: pix2pix stays in training mode, cyclegan in eval mode


def stage_i2i_time():
    """BASELINE configs[4]: unet_256 and resnet_9blocks (ngf 64) at 256x256 and 1024x1024, fp16, batch 1: device time of
    the engine's forward, and of the same module through torch's own CUDA ops (cuDNN) on the same GPU."""
    from innfer_b200 import synth
    torch.backends.cudnn.benchmark = True
    for family in ("unet", "resnet"):
        net = _i2i_net(family).to(dev).half()
        for size in (256, 1024):
            x = (torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev, torch.float16)
            flop = synth.i2i_flop(family, size, size)
            res = {}
            with torch.no_grad():
                for name, fn in (("ours", lambda t: net(t)), ("torch", lambda t: net.model(t.clone()))):
                    for _ in range(3):
                        y = fn(x)
                    torch.cuda.synchronize()
                    l0 = N.kernel_launches()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    iters = 20
                    e0.record()
                    for _ in range(iters):
                        y = fn(x)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / iters
                    res[name] = (ms, y.float())
                    print("i2i %s %dx%d fp16 %-5s: %8.3f ms  %7.1f TFLOP/s  (%d launches/forward)" %
                          (family, size, size, name, ms, flop / ms / 1e9, (N.kernel_launches() - l0) // iters))
            d = (res["ours"][1] - res["torch"][1]).abs().max().item()
            print("i2i %s %dx%d: max |ours - torch fp16| = %.4f, speed-up over torch/cuDNN %.2fx" %
                  (family, size, size, d, res["torch"][0] / res["ours"][0]))
        net.invalidate_engine()
        del net
    return True


def stage_i2i_prof():
    """One forward of each generator at 1024x1024 (after a warm-up at the same size) for ncu launch lists."""
    for family in ("unet", "resnet"):
        net = _i2i_net(family).to(dev).half()
        x = (torch.rand(1, 3, 1024, 1024, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev, torch.float16)
        with torch.no_grad():
            net(x)
            torch.cuda.synchronize()
            net(x)
        torch.cuda.synchronize()
        net.invalidate_engine()
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True)
    a = ap.parse_args()
    t0 = time.time()
    ok = {"conv1": stage_conv1, "convs": stage_convs, "net": stage_net, "time": stage_time, "prof": stage_prof, "trace": stage_trace, "trace_pair": stage_trace_pair, "trace_up": stage_trace_up, "ppon": stage_ppon, "ppon_prof": stage_ppon_prof, "ppon_time": stage_ppon_time, "pan_time": stage_pan_time, "srres_time": stage_srres_time, "lat": stage_lat, "pan_prof": stage_pan_prof, "steady": stage_steady, "pix": stage_pix, "cfg3": stage_cfg3, "i2i_time": stage_i2i_time, "i2i_prof": stage_i2i_prof}[a.stage]()
    print("STAGE %s %s (%.1fs)" % (a.stage, "OK" if ok else "FAILED", time.time() - t0))
    sys.exit(0 if ok else 1)
