"""Colour helpers -- mirror of the reference's utils/colors.py (bgr_to_rgb 5-11, rgb_to_bgr 14-16,
bgra/rgba 19-26, srgb2linear 29-46, linear2srgb 49-60).  These host versions serve the explicit
``-cpu`` mode; on the CUDA path the same maths runs inside the pixel / colour-fix kernels."""
import numpy as np
import torch


def bgr_to_rgb(image: torch.Tensor) -> torch.Tensor:
    return image.flip(-3)


def rgb_to_bgr(image: torch.Tensor) -> torch.Tensor:
    return image.flip(-3)


def bgra_to_rgba(image: torch.Tensor) -> torch.Tensor:
    return image[[2, 1, 0, 3], :, :]


def rgba_to_bgra(image: torch.Tensor) -> torch.Tensor:
    return image[[2, 1, 0, 3], :, :]


def srgb2linear(srgb, gamma=2.4, th=0.04045):
    """uint8 sRGB image -> float32 linear RGB in [0, 1]."""
    a, att = 0.055, 12.92
    v = np.float32(srgb) / 255.0
    return np.where(v <= th, v / att, np.power((v + a) / (1 + a), gamma))


def linear2srgb(linear, gamma=2.4, th=0.0031308):
    """float32 linear RGB -> uint8 sRGB (values are truncated, not rounded, like the reference)."""
    a, att = 0.055, 12.92
    v = np.clip(linear.copy(), 0.0, 1.0)
    v = np.where(v <= th, v * att, (1 + a) * np.power(v, 1.0 / gamma) - a)
    return np.clip(v * 255.0, 0.0, 255).astype(np.uint8)
