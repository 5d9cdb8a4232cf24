"""Python handle over the native RRDB engine (include/innfer_b200.h).

torch is plumbing here: it owns device memory and streams; every FLOP of the CUDA path runs in
libinnfer_b200.so.  Nothing in this module falls back to torch ops on failure.
"""
import ctypes

import numpy as np
import torch

from . import _native as N


def _dtype_code(dtype):
    if dtype == torch.float16:
        return N.INNFER_F16
    if dtype == torch.float32:
        return N.INNFER_F32
    raise TypeError("innfer_b200: tensors must be float16 or float32, got %s" % dtype)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class RRDBEngine:
    """One native network handle (weights repacked for the sm_100a kernels) on one CUDA device."""

    def __init__(self, cfg, device, fp16=True):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("RRDBEngine needs a CUDA device; the -cpu mode runs the torch modules instead")
        if not torch.cuda.is_available():
            raise RuntimeError("innfer_b200: CUDA is not available and there is no CPU fallback for the engine")
        self.lib = N.load()
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.fp16 = bool(fp16)
        self.cfg = dict(cfg)
        self._h = ctypes.c_void_p()
        self._create()
        self._finalized = False

    def _create(self):
        cfg = self.cfg
        c = N.RRDBCfg(cfg["in_nc"], cfg["out_nc"], cfg["nf"], cfg["nb"], cfg.get("gc", 32), cfg["scale"],
                      int(bool(cfg.get("plus", False))), int(self.fp16))
        N.check(self.lib.innfer_rrdb_create(ctypes.byref(c), self.index, ctypes.byref(self._h)))

    # -- construction ----------------------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, sd, cfg, device, fp16=True):
        eng = cls(cfg, device, fp16)
        try:
            for key, val in sd.items():
                eng.load(key, val)
            eng.finalize()
        except Exception:
            eng.close()
            raise
        return eng

    @classmethod
    def from_module(cls, module, device, fp16=True):
        """Build from an architectures.RRDBNet_arch.RRDBNet (uses its reference-named state dict)."""
        return cls.from_state_dict(module.state_dict(), module.cfg, device, fp16)

    def load(self, key, tensor):
        a = np.ascontiguousarray(tensor.detach().to("cpu", torch.float32).numpy())
        shape = (ctypes.c_int64 * a.ndim)(*a.shape)
        N.check(self.lib.innfer_rrdb_load(self._h, key.encode(), a.ctypes.data_as(ctypes.c_void_p), shape, a.ndim))

    def finalize(self):
        N.check(self.lib.innfer_rrdb_finalize(self._h))
        self._finalized = True

    def set_max_batch(self, n):
        N.check(self.lib.innfer_rrdb_set_max_batch(self._h, int(n)))

    def profile_reset(self, enable=True):
        """Start (1 / True), start per-launch family timing (2) or stop (0) device-side timing of the conv
        sequence; see innfer_rrdb_profile."""
        N.check(self.lib.innfer_rrdb_profile(self._h, int(enable)))

    def profile_read(self):
        """(milliseconds spent in the conv sequences, conv kernels launched) since profile_reset."""
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        N.check(self.lib.innfer_rrdb_profile_read(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, int(n.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.innfer_rrdb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- execution -------------------------------------------------------------------------------
    def _check_input(self, x):
        if x.dim() != 4 or x.shape[1] != self.cfg["in_nc"]:
            raise ValueError("expected [N,%d,H,W] input, got %s" % (self.cfg["in_nc"], tuple(x.shape)))
        if not x.is_cuda or (x.device.index is not None and x.device.index != self.index):
            raise ValueError("input must live on %s" % self.device)
        return x.contiguous()

    def forward(self, x):
        """RRDBNet.forward on a batch: [N,in_nc,h,w] -> [N,out_nc,s*h,s*w], same dtype."""
        x = self._check_input(x)
        n, _, h, w = x.shape
        s = self.cfg["scale"]
        y = torch.empty((n, self.cfg["out_nc"], s * h, s * w), dtype=x.dtype, device=x.device)
        with torch.cuda.device(self.index):
            N.check(self.lib.innfer_rrdb_forward(self._h, x.data_ptr(), n, h, w, y.data_ptr(), _dtype_code(x.dtype),
                                                 _stream_ptr(x.device)))
        return y

    def chop_forward(self, x, patch_size=200, step=0.5):
        """Model.chop_forward (tile, forward, blend) on a [1,in_nc,H,W] image in one native call."""
        x = self._check_input(x)
        if x.shape[0] != 1:
            raise ValueError("chop_forward expects batch size 1 (run.py:178-181 squeezes the batch)")
        _, _, H, W = x.shape
        s = self.cfg["scale"]
        y = torch.empty((1, self.cfg["out_nc"], s * H, s * W), dtype=x.dtype, device=x.device)
        with torch.cuda.device(self.index):
            N.check(self.lib.innfer_rrdb_chop_forward(self._h, x.data_ptr(), H, W, int(patch_size), float(step),
                                                      y.data_ptr(), _dtype_code(x.dtype), _stream_ptr(x.device)))
        return y

    def _check_u8_out(self, out, H, W, device):
        """A caller-supplied uint8 [s*H, s*W, 3] result buffer: host (numpy / CPU tensor) or on `device`."""
        s = self.cfg["scale"]
        shape = (s * H, s * W, 3)
        if isinstance(out, torch.Tensor):
            ok = out.dtype == torch.uint8 and tuple(out.shape) == shape and out.is_contiguous()
            on_dev = out.is_cuda and (out.device.index is None or out.device.index == self.index)
            ok = ok and (on_dev if device else not out.is_cuda)
        else:
            ok = (not device and isinstance(out, np.ndarray) and out.dtype == np.uint8 and out.shape == shape
                  and out.flags["C_CONTIGUOUS"] and out.flags["WRITEABLE"])
        if not ok:
            raise ValueError("`out` must be a contiguous uint8 %s %s" % (shape, "tensor on %s" % self.device if device
                                                                        else "host array"))
        return out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data

    def upscale_u8(self, img, patch_size=200, step=0.5, out=None):
        """np2tensor -> chop_forward -> tensor2np fused.  ``img``: HOST uint8 HWC BGR array (numpy, or
        a pinned CPU torch tensor); returns a HOST uint8 array [s*H, s*W, 3].  H2D and D2H copies are
        part of the call."""
        if isinstance(img, torch.Tensor):
            if img.dtype != torch.uint8 or img.is_cuda or not img.is_contiguous() or img.dim() != 3:
                raise ValueError("expected a contiguous CPU uint8 HWC tensor")
            src_ptr, (H, W, C) = img.data_ptr(), img.shape
        else:
            img = np.ascontiguousarray(img)
            if img.dtype != np.uint8 or img.ndim != 3:
                raise ValueError("expected a uint8 HWC image")
            src_ptr, (H, W, C) = img.ctypes.data, img.shape
        if C != 3:
            raise ValueError("the uint8 path handles 3-channel images")
        s = self.cfg["scale"]
        if out is None:
            out = np.empty((s * H, s * W, 3), dtype=np.uint8)
        dst_ptr = self._check_u8_out(out, H, W, device=False)
        with torch.cuda.device(self.index):
            N.check(self.lib.innfer_rrdb_upscale_u8(self._h, src_ptr, H, W, int(patch_size), float(step), dst_ptr,
                                                    _stream_ptr(self.device)))
        return out

    def _check_u8_device_image(self, img):
        if not isinstance(img, torch.Tensor) or img.dtype != torch.uint8 or not img.is_cuda or img.dim() != 3 \
                or img.shape[2] != 3 or not img.is_contiguous() or \
                (img.device.index is not None and img.device.index != self.index):
            raise ValueError("expected a contiguous uint8 [H, W, 3] tensor on %s" % self.device)
        return img.shape[0], img.shape[1]

    def upscale_u8_device(self, img, patch_size=200, step=0.5, out=None):
        """Same with DEVICE uint8 tensors (no copies, no synchronisation)."""
        H, W = self._check_u8_device_image(img)
        s = self.cfg["scale"]
        if out is None:
            out = torch.empty((s * H, s * W, 3), dtype=torch.uint8, device=img.device)
        dst = self._check_u8_out(out, H, W, device=True)
        with torch.cuda.device(self.index):
            N.check(self.lib.innfer_rrdb_upscale_u8_device(self._h, img.data_ptr(), H, W, int(patch_size), float(step),
                                                           dst, _stream_ptr(img.device)))
        return out

    def chop_forward_ex(self, x, patch_size=200, step=0.5, out_u8=False):
        """chop_forward with independent element types at the two ends (innfer_rrdb_chop_forward_ex): ``x`` is a
        device uint8 [H, W, 3] BGR image (np2tensor fused) or a [1, C, H, W] fp16/fp32 tensor; the result is a
        device uint8 [s*H, s*W, 3] BGR image (tensor2np fused) if ``out_u8`` else a [1, C, s*H, s*W] tensor of the
        engine's precision.  Lets a model chain stay on the device with float tensors between the models."""
        s = self.cfg["scale"]
        if x.dtype == torch.uint8:
            H, W = self._check_u8_device_image(x)
            xcode = N.INNFER_U8
        else:
            x = self._check_input(x)
            if x.shape[0] != 1:
                raise ValueError("chop_forward expects batch size 1")
            H, W = x.shape[2], x.shape[3]
            xcode = _dtype_code(x.dtype)
        if out_u8:
            y = torch.empty((s * H, s * W, 3), dtype=torch.uint8, device=x.device)
            ycode = N.INNFER_U8
        else:
            dt = torch.float16 if self.fp16 else torch.float32
            y = torch.empty((1, self.cfg["out_nc"], s * H, s * W), dtype=dt, device=x.device)
            ycode = _dtype_code(dt)
        with torch.cuda.device(self.index):
            N.check(self.lib.innfer_rrdb_chop_forward_ex(self._h, x.data_ptr(), xcode, H, W, int(patch_size), float(step),
                                                         y.data_ptr(), ycode, _stream_ptr(x.device)))
        return y

    def profile_families(self):
        """Per kernel family since profile_reset(2): {name: (launches, total ms, algorithmic FLOP, algorithmic bytes)}."""
        need = ctypes.c_uint64()
        N.check(self.lib.innfer_rrdb_profile_families(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        N.check(self.lib.innfer_rrdb_profile_families(self._h, buf, need.value, ctypes.byref(need)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms, flop, nbytes = line.split("\t")
            out[name] = (int(n), float(ms), float(flop), float(nbytes))
        return out


class SRResNetEngine(RRDBEngine):
    """Native handle for architectures.SRResNet_arch.SRResNet (same execution API as RRDBEngine)."""

    def _create(self):
        cfg = self.cfg
        mode = {"pixelshuffle": 0, "upconv": 1}[cfg.get("upsample_mode", "pixelshuffle")]
        c = N.SRResNetCfg(cfg["in_nc"], cfg["out_nc"], cfg["nf"], cfg["nb"], cfg["scale"], mode,
                          float(cfg.get("res_scale", 1.0)), int(self.fp16))
        N.check(self.lib.innfer_srresnet_create(ctypes.byref(c), self.index, ctypes.byref(self._h)))


class PPONEngine(RRDBEngine):
    """Native handle for architectures.PPON_arch.PPON; ``forward`` / ``chop_forward`` return out_p."""

    def _create(self):
        cfg = self.cfg
        c = N.PPONCfg(cfg["in_nc"], cfg["out_nc"], cfg["nf"], cfg["nb"], cfg["scale"], float(cfg.get("alpha", 1.0)),
                      int(self.fp16))
        N.check(self.lib.innfer_ppon_create(ctypes.byref(c), self.index, ctypes.byref(self._h)))


class PANEngine(RRDBEngine):
    """Native handle for architectures.PAN_arch.PAN (same execution API as RRDBEngine)."""

    def _create(self):
        cfg = self.cfg
        c = N.PANCfg(cfg["in_nc"], cfg["out_nc"], cfg["nf"], cfg["unf"], cfg["nb"], cfg["scale"],
                     int(bool(cfg.get("self_attention", True))), int(bool(cfg.get("double_scpa", False))), int(self.fp16))
        N.check(self.lib.innfer_pan_create(ctypes.byref(c), self.index, ctypes.byref(self._h)))


class _I2IEngine(RRDBEngine):
    """Native handle for the image-to-image generators (scale 1).  ``cfg`` carries ``train``: BatchNorm2d in training
    mode (run.py:297 keeps pix2pix that way) normalises with the statistics of the batch."""

    _kind = 0
    _depth_key = "num_downs"

    @classmethod
    def from_module(cls, module, device, fp16=True, unit_io=False):
        cfg = dict(module.cfg, train=bool(module.training), unit_io=bool(unit_io))
        return cls.from_state_dict(module.state_dict(), cfg, device, fp16)

    def _create(self):
        cfg = self.cfg
        c = N.I2ICfg(self._kind, cfg["in_nc"], cfg["out_nc"], cfg["ngf"], cfg[self._depth_key],
                     {"batch": 0, "instance": 1}[cfg["norm"]], int(bool(cfg.get("train", False))), int(self.fp16),
                     int(bool(cfg.get("unit_io", False))))
        N.check(self.lib.innfer_i2i_create(ctypes.byref(c), self.index, ctypes.byref(self._h)))

    def load(self, key, tensor):
        if tensor.dim() == 0:      # BatchNorm2d.num_batches_tracked: bookkeeping, not arithmetic
            return
        super().load(key, tensor)


class UNetEngine(_I2IEngine):
    """architectures.UNet_arch.UnetGenerator (pix2pix)."""


class ResNetGenEngine(_I2IEngine):
    """architectures.ResNet_arch.ResnetGenerator (CycleGAN)."""

    _kind = 1
    _depth_key = "n_blocks"
