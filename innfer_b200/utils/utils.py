"""Image / tensor helpers -- mirror of the hot-path subset of the reference's utils/utils.py:
file scanning (36-65), read_img / save_img (68-133), np2tensor (164-194), tensor2np (197-248),
color_fix (278-315), extract_patches_2d (318-369), recompose_tensor (372-445), mod2normal (666-698),
swa2normal (701-720).

Tensor helpers accept CPU tensors (explicit ``-cpu`` mode, plain torch) and CUDA tensors (native
kernels through the C-ABI; an error is raised if the library is missing -- never a silent
fallback).
"""
import os.path as osp
from collections import OrderedDict
from os import walk as osw

import cv2
import numpy as np
import torch

from .colors import bgr_to_rgb, bgra_to_rgba, linear2srgb, rgb_to_bgr, rgba_to_bgra, srgb2linear  # noqa: F401

MODEL_EXTENSIONS = [".pth", ".pt"]
IMG_EXTENSIONS = [".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".webp", "tga", ".tif", ".tiff", ".dng"]

_MAXVAL = {np.dtype("uint8"): 255, np.dtype("uint16"): 65535, np.dtype("int8"): 127, np.dtype("int16"): 32767,
           np.dtype("float32"): 1.0, np.dtype("float64"): 1.0}


# ------------------------------------------------------------------------------------ files
def is_ext_file(filename, extensions=IMG_EXTENSIONS):
    return any(filename.endswith(ext) for ext in extensions)


def scan_dir(path, extensions=IMG_EXTENSIONS):
    if not osp.isdir(path):
        raise AssertionError(f"{path:s} is not a valid directory")
    found = []
    for dirpath, _, fnames in sorted(osw(path)):
        found.extend(osp.join(dirpath, f) for f in sorted(fnames) if is_ext_file(f, extensions))
    return found


def get_models_paths(path):
    models = scan_dir(path, MODEL_EXTENSIONS)
    if not models:
        raise AssertionError(f"{path:s} has no valid model file")
    return models


def get_images_paths(path):
    images = scan_dir(path, IMG_EXTENSIONS)
    if not images:
        raise AssertionError(f"{path:s} has no valid image file")
    return images


def read_img(path=None):
    """cv2.imread(IMREAD_UNCHANGED): HWC BGR uint8 (or None when unreadable). Gray images get a
    channel axis.  (.dng/rawpy input of the reference is not part of the hot path.)"""
    if path is None or path.lower().endswith(".dng"):
        return None
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is not None and img.ndim == 2:
        img = np.expand_dims(img, axis=2)
    return img


def save_img(img, img_path, mode="RGB"):
    cv2.imwrite(img_path, img)


def merge_imgs(img_list):
    """Side-by-side comparison strip: every image nearest-resized to the largest height."""
    h = max(i.shape[0] for i in img_list)
    w = max(i.shape[1] for i in img_list)
    return cv2.hconcat([cv2.resize(i, (w, h), interpolation=cv2.INTER_NEAREST) if i.shape[:2] != (h, w) else i
                        for i in img_list])


def save_img_comp(img_list, img_path, mode="RGB"):
    save_img(merge_imgs(img_list), img_path, mode)


# ------------------------------------------------------------------------------------ tensors
def norm(x):
    out = (x - 0.5) * 2.0
    return out.clamp(-1, 1) if isinstance(x, torch.Tensor) else np.clip(out, -1, 1)


def denorm(x, min_max=(-1.0, 1.0)):
    out = (x - min_max[0]) / (min_max[1] - min_max[0])
    return out.clamp(0, 1) if isinstance(x, torch.Tensor) else np.clip(out, 0, 1)


def np2tensor(img, bgr2rgb=True, data_range=1.0, normalize=False, change_range=True, add_batch=True):
    """uint8 HWC (BGR) numpy image -> float32 [1,C,H,W] (RGB) tensor in [0,1]."""
    if not isinstance(img, np.ndarray):
        raise TypeError("Got unexpected object type, expected np.ndarray")
    if change_range:
        img = img.astype(np.float32) / _MAXVAL.get(img.dtype, 1.0)
    t = torch.from_numpy(np.ascontiguousarray(np.transpose(img, (2, 0, 1)))).float()
    if bgr2rgb:
        if t.shape[0] % 3 == 0:
            t = bgr_to_rgb(t)
        elif t.shape[0] == 4:
            t = bgra_to_rgba(t)
    if add_batch:
        t = t.unsqueeze(0)
    if normalize:
        t = norm(t)
    return t


def tensor2np(img, rgb2bgr=True, remove_batch=True, data_range=255, denormalize=False, change_range=True,
              imtype=np.uint8):
    """[1,C,H,W] / [C,H,W] / [H,W] tensor (RGB) -> HWC (BGR) numpy image, clip(255x).round()."""
    if not isinstance(img, torch.Tensor):
        raise TypeError("Got unexpected object type, expected torch.Tensor")
    n_dim = img.dim()
    img = img.float().cpu()
    if n_dim in (3, 4):
        if n_dim == 4 and remove_batch:
            img = img.squeeze(dim=0)
        if img.shape[0] == 3 and rgb2bgr:
            arr = rgb_to_bgr(img).numpy()
        elif img.shape[0] == 4 and rgb2bgr:
            arr = rgba_to_bgra(img).numpy()
        else:
            arr = img.numpy()
        arr = np.transpose(arr, (1, 2, 0))
    elif n_dim == 2:
        arr = img.numpy()
    else:
        raise TypeError(f"Only support 4D, 3D and 2D tensor. But received with dimension: {n_dim:d}")
    if denormalize:
        arr = denorm(arr)
    if change_range:
        arr = np.clip(data_range * arr, 0, data_range).round()
    return arr.astype(imtype)


def modcrop(img_in, scale):
    img = np.copy(img_in)
    if img.ndim not in (2, 3):
        raise ValueError("Wrong img ndim: [{:d}].".format(img.ndim))
    h, w = img.shape[:2]
    return img[:h - h % scale, :w - w % scale]


def linear_resize(img, st=256):
    """Resize (bicubic, in linear light) up to the next multiple of ``st`` in both dimensions, as run.py does for the
    pix2pix UNets whose depth fixes the input granularity (utils.py:267-275 of the reference)."""
    h, w = img.shape[0:2]
    if h % st or w % st:
        oh, ow = -(-h // st) * st, -(-w // st) * st
        img = linear2srgb(cv2.resize(srgb2linear(img), dsize=(ow, oh), interpolation=cv2.INTER_CUBIC))
    return img


# ------------------------------------------------------------------------------------ colour fix
def color_fix(imgA, imgB, device=None):
    """Add the low-frequency difference (LR - downscaled SR) back to SR in linear light.

    ``device`` None/cpu: host implementation (cv2 + numpy, as the reference).  A CUDA device: the
    three native kernels (innfer_color_fix); failures raise."""
    if device is not None and torch.device(device).type == "cuda":
        from .. import _native as N
        lib = N.load()
        a = np.ascontiguousarray(imgA)
        b = np.ascontiguousarray(imgB)
        if a.dtype != np.uint8 or b.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3 or b.shape[2] != 3:
            raise ValueError("color_fix on CUDA expects uint8 HWC 3-channel images")
        out = np.empty_like(b)
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        N.check(lib.innfer_color_fix_host(a.ctypes.data, a.shape[0], a.shape[1], b.ctypes.data, b.shape[0],
                                          b.shape[1], out.ctypes.data, idx))
        return out
    a = srgb2linear(imgA)
    b = srgb2linear(imgB)
    ha, wa = a.shape[:2]
    hb, wb = b.shape[:2]
    scaling = ha < hb and wa < wb
    b_small = cv2.resize(b, dsize=(wa, ha), interpolation=cv2.INTER_CUBIC) if scaling else b
    low = cv2.GaussianBlur(a - b_small, (3, 3), 0)
    if scaling:
        low = cv2.resize(low, dsize=(wb, hb), interpolation=cv2.INTER_CUBIC)
    return linear2srgb(low + b)


def color_fix_device(lr, sr, out=None):
    """color_fix on DEVICE uint8 HWC tensors (lr [h,w,3], sr [H,W,3]) -> device uint8 [H,W,3]; the three native
    kernels of innfer_color_fix on the current stream, no host round trip."""
    from .. import _native as N
    for t in (lr, sr):
        if not isinstance(t, torch.Tensor) or t.dtype != torch.uint8 or not t.is_cuda or t.dim() != 3 or t.shape[2] != 3 \
                or not t.is_contiguous():
            raise ValueError("color_fix_device expects contiguous CUDA uint8 [H, W, 3] tensors")
    if lr.device != sr.device:
        raise ValueError("both images must live on the same device")
    if out is None:
        out = torch.empty_like(sr)
    elif out.dtype != torch.uint8 or out.shape != sr.shape or out.device != sr.device or not out.is_contiguous():
        raise ValueError("`out` must match the SR image")
    with torch.cuda.device(sr.device):
        N.check(N.load().innfer_color_fix(lr.data_ptr(), lr.shape[0], lr.shape[1], sr.data_ptr(), sr.shape[0], sr.shape[1],
                                          out.data_ptr(), torch.cuda.current_stream(sr.device).cuda_stream))
    return out


# ------------------------------------------------------------------------------------ tiling
def _tile_starts(length, size, stride):
    starts = list(range(0, length - size + 1, stride))
    if (length - size) % stride != 0:
        starts.append(length - size)
    return starts


def extract_patches_2d(img, patch_shape, step=None, batch_first=False):
    """[B,C,H,W] -> [nP,B,C,ph,pw] (or [B,nP,...]) sliding crops; stride = int(patch * step) for
    float steps; an extra crop is anchored to the far edge when the stride does not divide."""
    if step is None:
        step = [1.0, 1.0]
    ph, pw = patch_shape
    if img.size(2) < ph or img.size(3) < pw:
        pad_h, pad_w = max(ph - img.size(2), 0), max(pw - img.size(3), 0)
        img = torch.nn.functional.pad(img, (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2))
    sh = int(ph * step[0]) if isinstance(step[0], float) else step[0]
    sw = int(pw * step[1]) if isinstance(step[1], float) else step[1]
    ys = _tile_starts(img.size(2), ph, sh)
    xs = _tile_starts(img.size(3), pw, sw)
    patches = torch.stack([img[:, :, y:y + ph, x:x + pw] for y in ys for x in xs], 0)
    return patches.permute(1, 0, 2, 3, 4) if batch_first else patches


def recompose_tensor(patches, height, width, step=None, scale=1):
    """Blend [nP,C,P,P] tiles back into [B,C,scale*H,scale*W] with the 0.1->1.0 linear cross-fade.
    CUDA tiles go through the native gather-blend kernel, CPU tiles through plain torch."""
    if step is None:
        step = [1.0, 1.0]
    assert isinstance(step, float) and 0.5 <= step <= 1.0
    full_h, full_w = scale * height, scale * width
    n, ch, P, _ = patches.size()
    if patches.is_cuda:
        return _recompose_native(patches, height, width, step, scale)
    overlap = scale * int(round((1.0 - step) * (P / scale)))
    eff = int(step * P)
    stride = int(P * step)
    rows = 1 + (max(full_h, P) - P) // stride + (1 if (max(full_h, P) - P) % stride else 0)
    cols = 1 + (max(full_w, P) - P) // stride + (1 if (max(full_w, P) - P) % stride else 0)
    batch = n // (rows * cols)
    ramp_up = torch.linspace(0.1, 1.0, overlap, dtype=patches.dtype)
    ramp_dn = torch.linspace(1.0, 0.1, overlap, dtype=patches.dtype)
    prof = torch.cat([ramp_up, torch.ones(P - 2 * overlap, dtype=patches.dtype), ramp_dn], 0)
    wpatch = prof[None, :] * prof[:, None]
    wsum = torch.zeros(1, ch, full_h, full_w, dtype=patches.dtype)
    out = torch.zeros(batch, ch, full_h, full_w, dtype=patches.dtype)
    i = 0
    for b in range(batch):
        for r in range(rows):
            for c in range(cols):
                y0, x0 = min(r * eff, full_h - P), min(c * eff, full_w - P)
                if b == 0:
                    wsum[0, :, y0:y0 + P, x0:x0 + P] += wpatch
                out[b, :, y0:y0 + P, x0:x0 + P] += patches[i] * wpatch
                i += 1
    return out / wsum


def _recompose_native(patches, height, width, step, scale):
    """CUDA branch of recompose_tensor: the gather-blend kernel, in the tiles' own precision (fp16 tiles blend from
    fp16, fp32 tiles from fp32), one image at a time for a batch (final_batch_size = n // tiles per image,
    utils.py:412 of the reference)."""
    import ctypes

    from .. import _native as N
    lib = N.load()
    n, ch, P, Pw = patches.shape
    if P != Pw:
        raise ValueError("square tiles expected, got %dx%d" % (P, Pw))
    if ch > 8:
        raise ValueError("native blend supports up to 8 channels")
    if patches.dtype not in (torch.float16, torch.float32):
        raise TypeError("native blend expects float16 or float32 tiles, got %s" % patches.dtype)
    if P % scale:
        raise ValueError("tile size %d is not a multiple of the scale %d" % (P, scale))
    p = P // scale
    nt, ts = ctypes.c_int(), ctypes.c_int()
    N.check(lib.innfer_tiles_plan(height, width, p, float(step), None, 0, ctypes.byref(nt), ctypes.byref(ts)))
    if ts.value != p or nt.value < 1 or n % nt.value:
        raise ValueError("%d tiles of %d px do not cover a %dx%d image (%d tiles of %d px per image expected)"
                         % (n, p, height, width, nt.value, ts.value))
    batch = n // nt.value
    f16 = patches.dtype == torch.float16
    # tiles -> planar-chunk [n][1][P][P][8]
    chunks = torch.zeros(n, 1, P, P, 8, dtype=patches.dtype, device=patches.device)
    chunks[:, 0, :, :, :ch] = patches.permute(0, 2, 3, 1)
    out = torch.empty(batch, ch, scale * height, scale * width, dtype=patches.dtype, device=patches.device)
    code = N.INNFER_F16 if f16 else N.INNFER_F32
    fn = lib.innfer_blend if f16 else lib.innfer_blend_f32
    stream = torch.cuda.current_stream(patches.device).cuda_stream
    with torch.cuda.device(patches.device):
        for b in range(batch):
            N.check(fn(chunks[b * nt.value:(b + 1) * nt.value].data_ptr(), height, width, p, float(step), scale, ch,
                       out[b].data_ptr(), code, stream))
    return out


# ------------------------------------------------------------------------------------ state dicts
def mod2normal(state_dict):
    """'modified/new-arch' ESRGAN keys (conv_first, RRDB_trunk, ...) -> original keys. Like the
    reference this assumes 23 blocks and a 4x tail."""
    if "conv_first.weight" not in state_dict:
        return state_dict
    print("Converting and loading a modified RRDB model to normal RRDB")
    out = OrderedDict()
    out["model.0.weight"] = state_dict["conv_first.weight"]
    out["model.0.bias"] = state_dict["conv_first.bias"]
    for key, val in state_dict.items():
        if "RDB" not in key:
            continue
        nk = key.replace("RRDB_trunk.", "model.1.sub.")
        if ".weight" in key:
            nk = nk.replace(".weight", ".0.weight")
        elif ".bias" in key:
            nk = nk.replace(".bias", ".0.bias")
        out[nk] = val
    fixed = (("model.1.sub.23", "trunk_conv"), ("model.3", "upconv1"), ("model.6", "upconv2"),
             ("model.8", "HRconv"), ("model.10", "conv_last"))
    for new, old in fixed:
        out[new + ".weight"] = state_dict[old + ".weight"]
        out[new + ".bias"] = state_dict[old + ".bias"]
    return out


def normal2mod(state_dict):
    """Inverse of mod2normal (utils.py:629-663); not used by run.py, kept for API parity."""
    if "model.0.weight" not in state_dict:
        return state_dict
    print("Converting and loading an RRDB model to modified RRDB")
    out = OrderedDict()
    out["conv_first.weight"] = state_dict["model.0.weight"]
    out["conv_first.bias"] = state_dict["model.0.bias"]
    for key, val in state_dict.items():
        if "RDB" not in key:
            continue
        nk = key.replace("model.1.sub.", "RRDB_trunk.")
        if ".0.weight" in key:
            nk = nk.replace(".0.weight", ".weight")
        elif ".0.bias" in key:
            nk = nk.replace(".0.bias", ".bias")
        out[nk] = val
    fixed = (("trunk_conv", "model.1.sub.23"), ("upconv1", "model.3"), ("upconv2", "model.6"),
             ("HRconv", "model.8"), ("conv_last", "model.10"))
    for new, old in fixed:
        out[new + ".weight"] = state_dict[old + ".weight"]
        out[new + ".bias"] = state_dict[old + ".bias"]
    return out


def swa2normal(state_dict):
    """Strip the torch SWA wrapper ('n_averaged' + 'module.module.' prefixes)."""
    if "n_averaged" not in state_dict:
        return state_dict
    print("Attempting to convert a SWA model to a regular model\n")
    out = OrderedDict()
    for key, val in state_dict.items():
        if "n_averaged" in key:
            print("n_averaged: {}".format(val))
        elif "module.module." in key:
            out[key.replace("module.module.", "")] = val
    return out
