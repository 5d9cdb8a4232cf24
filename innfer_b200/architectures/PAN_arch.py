"""PAN generator -- mirror of the reference's architectures/PAN_arch.py (PA 22-36, PACnv 38-57, SCPA 59-103,
PAN 104-222) and of block.SelfAttentionBlock (block.py:398-473), with the reference's parameter names, so that
``load_state_dict(strict)`` and run.py's key probe ('SCPA_trunk.0.conv1_a.weight') behave identically.

The module tree owns the parameters and runs the explicit ``-cpu`` mode; a CUDA tensor goes to the sm_100a engine
(SURVEY.md 8f rank 3).  Only the 'nearest' upsampler (the default, utils/defaults.py:89) is built.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import block as B
from ._native import NativeEngineMixin


class PA(nn.Module):
    """Pixel attention: x * sigmoid(conv1x1(x))."""

    def __init__(self, nf):
        super().__init__()
        self.conv = nn.Conv2d(nf, nf, 1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        return x * self.sigmoid(self.conv(x))


class PACnv(nn.Module):
    def __init__(self, nf, k_size=3):
        super().__init__()
        self.k2 = nn.Conv2d(nf, nf, 1)
        self.sigmoid = nn.Sigmoid()
        self.k3 = nn.Conv2d(nf, nf, k_size, padding=(k_size - 1) // 2, bias=False)
        self.k4 = nn.Conv2d(nf, nf, k_size, padding=(k_size - 1) // 2, bias=False)

    def forward(self, x):
        return self.k4(self.k3(x) * self.sigmoid(self.k2(x)))


class SCPA(nn.Module):
    """Self-calibrated block: two 1x1 branches of nf/2 channels, a plain 3x3 and a PACnv, 1x1 fusion, residual."""

    def __init__(self, nf, reduction=2, stride=1, dilation=1):
        super().__init__()
        gw = nf // reduction
        self.conv1_a = nn.Conv2d(nf, gw, 1, bias=False)
        self.conv1_b = nn.Conv2d(nf, gw, 1, bias=False)
        self.k1 = nn.Sequential(nn.Conv2d(gw, gw, 3, stride, dilation, dilation, bias=False))
        self.PACnv = PACnv(gw)
        self.conv3 = nn.Conv2d(gw * reduction, nf, 1, bias=False)
        self.lrelu = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        a = self.lrelu(self.k1(self.lrelu(self.conv1_a(x))))
        b = self.lrelu(self.PACnv(self.lrelu(self.conv1_b(x))))
        return self.conv3(torch.cat([a, b], 1)) + x


class SelfAttentionBlock(nn.Module):
    """SAGAN attention over the max-pooled map, bicubic resize back, gamma * out + input (max_pool variant only)."""

    def __init__(self, in_dim, poolsize=4):
        super().__init__()
        self.pooled = nn.MaxPool2d(poolsize, poolsize)
        self.conv_f = nn.Conv1d(in_dim, in_dim // 8, 1)
        self.conv_g = nn.Conv1d(in_dim, in_dim // 8, 1)
        self.conv_h = nn.Conv1d(in_dim, in_dim, 1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, inp):
        x = self.pooled(inp)
        n, c, hp, wp = x.shape
        x = x.view(n, c, hp * wp)
        att = torch.softmax(torch.bmm(self.conv_f(x).transpose(1, 2), self.conv_g(x)), dim=-1)
        out = torch.bmm(self.conv_h(x), att.transpose(1, 2)).view(n, c, hp, wp)
        out = F.interpolate(out, size=inp.shape[2:], mode="bicubic", align_corners=False)
        return self.gamma * out + inp


def pa_upconv_block(nf, unf, upscale_factor=2, mode="nearest"):
    """Upsample, conv, PA, LeakyReLU, conv, and the SAME LeakyReLU instance again: a flattening ``B.sequential``
    keeps it once (children() de-duplicates), which is how the reference behaves for scales 4 and 8."""
    a = B.act("lrelu")
    return B.sequential(B.Upsample(scale_factor=upscale_factor, mode=mode), nn.Conv2d(nf, unf, 3, 1, 1), PA(unf), a,
                        nn.Conv2d(unf, unf, 3, 1, 1), a)


class PAN(NativeEngineMixin, nn.Module):
    _engine_class = "PANEngine"

    def __init__(self, in_nc, out_nc, nf, unf, nb, scale=4, self_attention=True, double_scpa=False,
                 ups_inter_mode="nearest"):
        super().__init__()
        if ups_inter_mode != "nearest":
            raise NotImplementedError("PAN: only ups_inter_mode='nearest' is supported")
        n_upscale = 1 if scale == 3 else int(math.log(scale, 2))
        if scale == 1:
            unf = nf
        self.scale = scale
        self.self_attention = self_attention
        self.double_scpa = double_scpa
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, unf=unf, nb=nb, scale=scale,
                        self_attention=bool(self_attention), double_scpa=bool(double_scpa))
        # construction order = the reference's (the default initialisation consumes the RNG in this order)
        self.conv_first = nn.Conv2d(in_nc, nf, 3, 1, 1)
        self.SCPA_trunk = nn.Sequential(*[SCPA(nf) for _ in range(nb)])
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1)
        if double_scpa:
            self.SCPA_trunk2 = nn.Sequential(*[SCPA(nf) for _ in range(nb)])
            self.trunk_conv2 = nn.Conv2d(nf, nf, 3, 1, 1)
        if self_attention:
            self.FSA = SelfAttentionBlock(nf)
        stages = [pa_upconv_block(nf if i == 0 else unf, unf, 3 if scale == 3 else 2) for i in range(n_upscale)]
        self.upsample = B.sequential(*stages)
        self.conv_last = nn.Conv2d(unf, out_nc, 3, 1, 1)

    def forward(self, x):
        if x.is_cuda:
            return self._engine(x.device, x.dtype).forward(x)
        fea = self.conv_first(x)
        trunk = self.trunk_conv(self.SCPA_trunk(fea))
        if self.double_scpa:
            trunk = self.trunk_conv2(self.SCPA_trunk2(trunk))
        fea = fea + trunk
        if self.self_attention:
            fea = self.FSA(fea)
        out = self.conv_last(self.upsample(fea))
        if self.scale > 1:
            return out + F.interpolate(x, scale_factor=self.scale, mode="bilinear", align_corners=True)
        return out + x
