"""Building blocks of the RRDB path with the reference's module structure and parameter names.

Mirror of the hot-path subset of the reference's architectures/block.py: act (81-101),
get_valid_padding (163-166), ShortcutBlock (183-194), sequential (197-210), conv_block (213-254),
Upsample (286-331), upconv_block (348-361), conv_layer (364-366), conv1x1 (390-391), GaussianNoise (375-388).
These modules carry the parameters (so load_state_dict sees the reference key names) and execute
the explicit ``-cpu`` mode; on a CUDA device the owning RRDBNet bypasses them and runs the
sm_100a engine (innfer_b200.engine).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def act(act_type, inplace=True, neg_slope=0.2):
    kind = act_type.lower()
    if kind == "relu":
        return nn.ReLU(inplace)
    if kind in ("leakyrelu", "lrelu"):
        return nn.LeakyReLU(neg_slope, inplace)
    raise NotImplementedError("activation layer [%s] is not part of the RRDB hot path" % act_type)


def get_valid_padding(kernel_size, dilation):
    kernel_size = kernel_size + (kernel_size - 1) * (dilation - 1)
    return (kernel_size - 1) // 2


class ShortcutBlock(nn.Module):
    """x + sub(x)."""

    def __init__(self, submodule):
        super().__init__()
        self.sub = submodule

    def forward(self, x):
        return x + self.sub(x)


def sequential(*args):
    """nn.Sequential that drops None entries and splices nested Sequentials (keeps flat indices).
    A single argument is returned as is."""
    if len(args) == 1:
        return args[0]
    mods = []
    for m in args:
        if isinstance(m, nn.Sequential):
            mods.extend(m.children())
        elif isinstance(m, nn.Module):
            mods.append(m)
    return nn.Sequential(*mods)


def conv_block(in_nc, out_nc, kernel_size, stride=1, dilation=1, groups=1, bias=True, pad_type="zero",
               norm_type=None, act_type="relu", mode="CNA", convtype="Conv2D", spectral_norm=False):
    """Conv2d(+zero padding) [+ activation]; only the configuration the RRDB path uses is built."""
    if mode != "CNA" or norm_type or pad_type != "zero" or convtype != "Conv2D" or spectral_norm:
        raise NotImplementedError("conv_block: only mode='CNA', zero padding, no norm, Conv2D is supported")
    padding = get_valid_padding(kernel_size, dilation)
    conv = nn.Conv2d(in_nc, out_nc, kernel_size=kernel_size, stride=stride, padding=padding,
                     dilation=dilation, bias=bias, groups=groups)
    a = act(act_type) if act_type else None
    return sequential(None, conv, None, a)


class Upsample(nn.Module):
    def __init__(self, size=None, scale_factor=None, mode="nearest", align_corners=None):
        super().__init__()
        self.scale_factor = float(scale_factor) if scale_factor else None
        self.mode = mode
        self.size = size
        self.align_corners = align_corners

    def forward(self, x):
        return F.interpolate(x, size=self.size, scale_factor=self.scale_factor, mode=self.mode,
                             align_corners=self.align_corners)

    def extra_repr(self):
        return "scale_factor=%s, mode=%s" % (self.scale_factor, self.mode)


def upconv_block(in_nc, out_nc, upscale_factor=2, kernel_size=3, stride=1, bias=True, pad_type="zero",
                 norm_type=None, act_type="relu", mode="nearest", convtype="Conv2D"):
    up = Upsample(scale_factor=upscale_factor, mode=mode)
    conv = conv_block(in_nc, out_nc, kernel_size, stride, bias=bias, pad_type=pad_type,
                      norm_type=norm_type, act_type=act_type, convtype=convtype)
    return sequential(up, conv)


def pixelshuffle_block(in_nc, out_nc, upscale_factor=2, kernel_size=3, stride=1, bias=True, pad_type="zero",
                       norm_type=None, act_type="relu", convtype="Conv2D"):
    """conv (in_nc -> out_nc * r^2) + PixelShuffle(r) [+ activation] (reference block.py:333-346)."""
    if norm_type:
        raise NotImplementedError("pixelshuffle_block: norm layers are not supported")
    conv = conv_block(in_nc, out_nc * (upscale_factor ** 2), kernel_size, stride, bias=bias, pad_type=pad_type,
                      norm_type=None, act_type=None, convtype=convtype)
    a = act(act_type) if act_type else None
    return sequential(conv, nn.PixelShuffle(upscale_factor), None, a)


class GaussianNoise(nn.Module):
    """Identity in eval mode (the only mode inference uses)."""

    def __init__(self, sigma=0.1):
        super().__init__()
        self.sigma = sigma

    def forward(self, x):
        if self.training and self.sigma != 0:
            x = x + torch.randn_like(x) * (self.sigma * x)
        return x


def conv_layer(in_channels, out_channels, kernel_size, stride=1, dilation=1, groups=1):
    """Bare Conv2d with "same" zero padding for the given dilation (reference block.py:364-366)."""
    padding = int((kernel_size - 1) / 2) * dilation
    return nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding=padding, bias=True, dilation=dilation,
                     groups=groups)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)
