"""Synthetic weights and work accounting for benchmarks and demos (SURVEY.md 8d recipe).

``make_state_dict`` builds the network through the package's own factory with PyTorch's default
initialisation under a fixed seed and sets the last conv's bias to 0.5, which keeps ~60 % of the
output pixels un-clipped so uint8 comparisons are meaningful.  (tests/ check that this reproduces
the reference's initialisation bit for bit.)
"""
import torch

from .architectures import get_network
from .utils.defaults import get_network_G_config


def make_state_dict(scale=4, nb=23, nf=64, in_nc=3, out_nc=3, seed=0, last_bias=0.5):
    torch.manual_seed(seed)
    net = get_network(get_network_G_config({"type": "esrgan", "nb": nb, "nf": nf, "in_nc": in_nc, "out_nc": out_nc}, scale))
    sd = net.state_dict()
    if last_bias is not None:
        last = max((int(k.split(".")[1]) for k in sd if k.count(".") == 2), default=None)
        sd["model.%d.bias" % last].fill_(last_bias)
    return sd


def flop_per_lr_pixel(scale=4, nb=23, nf=64, in_nc=3, out_nc=3):
    """Algorithmic conv FLOPs per low-res pixel pushed through RRDBNet (real channels only):
    35 853 696 for the 4x / 23-block / nf=64 net."""
    mac = 9 * in_nc * nf
    rdb = 9 * (sum((nf + 32 * k) * 32 for k in range(4)) + (nf + 128) * nf)
    mac += nb * 3 * rdb + 9 * nf * nf
    n_up = {1: 0, 2: 1, 3: 1, 4: 2, 8: 3}[scale]
    f = 3 if scale == 3 else 2
    res = 1
    for _ in range(n_up):
        res *= f * f
        mac += res * 9 * nf * nf
    mac += res * 9 * nf * nf + res * 9 * nf * out_nc
    return 2 * mac


def bytes_per_lr_pixel(scale=4, nb=23, nf=64, in_nc=3, out_nc=3):
    """Algorithmic fp16 activation bytes per low-res pixel moved by the conv sequence when every conv
    reads its real input channels once and writes its output once (residual inputs read once each):
    the HBM traffic of the layer-by-layer schedule whenever a batch of tiles is larger than L2."""
    e = 2
    b = (in_nc + nf) * e                                              # fea_conv
    rdb = sum((nf + 32 * k + 32) * e for k in range(4)) + (nf + 128 + nf + nf) * e   # conv1..4, conv5 (+x residual)
    b += nb * (3 * rdb + nf * e)                                      # + RRDB-level residual read
    b += 3 * nf * e                                                   # LR_conv: in, shortcut, out
    n_up = {1: 0, 2: 1, 3: 1, 4: 2, 8: 3}[scale]
    f = 3 if scale == 3 else 2
    res = 1
    for _ in range(n_up):
        b += res * nf * e
        res *= f * f
        b += res * nf * e
    b += res * 2 * nf * e                                             # HR_conv0
    b += res * (nf + max(out_nc, 4)) * e                              # HR_conv1 (compact 4-channel tile pixels)
    return b


def i2i_flop(kind, H, W, ngf=64, depth=None, in_nc=3, out_nc=3):
    """Algorithmic conv FLOPs of one H x W image through UnetGenerator (kind 'unet', depth = num_downs, default 8) or
    ResnetGenerator (kind 'resnet', depth = n_blocks, default 9): 2 x MACs of every (transposed) convolution
    (architectures/UNet_arch.py:47-66, ResNet_arch.py:55-91)."""
    mac = 0
    if kind == "unet":
        D = 8 if depth is None else depth
        inner = [ngf * min(2 ** i, 8) for i in range(D)]
        for i in range(D):
            outer = out_nc if i == 0 else inner[i - 1]
            cin = in_nc if i == 0 else outer
            px = (H >> (i + 1)) * (W >> (i + 1))          # pixels at the bottom of level i
            mac += px * 16 * cin * inner[i]               # 4x4 stride-2 conv
            mac += px * 16 * (inner[i] if i == D - 1 else 2 * inner[i]) * outer   # 4x4 stride-2 transposed conv
        return 2 * mac
    nb = 9 if depth is None else depth
    px = H * W
    mac += px * 49 * in_nc * ngf + px * 49 * ngf * out_nc
    mac += (px // 4) * 9 * ngf * 2 * ngf + (px // 16) * 9 * 2 * ngf * 4 * ngf
    mac += nb * 2 * (px // 16) * 9 * 4 * ngf * 4 * ngf
    mac += (px // 16) * 9 * 4 * ngf * 2 * ngf + (px // 4) * 9 * 2 * ngf * ngf
    return 2 * mac
