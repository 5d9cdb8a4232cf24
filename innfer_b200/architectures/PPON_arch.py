"""PPON generator -- mirror of the reference's architectures/PPON_arch.py:12-127 (PPON 12-76, _ResBlock_32 78-114,
RRBlock_32 116-127) with the reference's parameter names (CFEM / SFEM / PFEM / CRM / SRM / PRM), so that
``load_state_dict(strict)`` and the key-based architecture probe of run.py behave identically.

The module tree owns the parameters and runs the explicit ``-cpu`` mode.  A CUDA tensor goes to the sm_100a
engine, which computes what run.py consumes: the third output ``out_p`` (run.py:191-192,
220-221); on a CUDA device ``forward`` therefore returns ``(None, None, out_p)``.
"""
import math

import torch
import torch.nn as nn

from . import block as B
from ._native import NativeEngineMixin


class _ResBlock_32(nn.Module):
    """c1, eight 64->32 convs with dilation 1..8 whose running sums are concatenated, 1x1 fusion, 0.2 residual."""

    def __init__(self, nc=64):
        super().__init__()
        self.c1 = B.conv_layer(nc, nc, 3, 1, 1)
        for rate in range(1, 9):
            setattr(self, "d%d" % rate, B.conv_layer(nc, nc // 2, 3, 1, rate))
        self.act = B.act("lrelu")
        self.c2 = B.conv_layer(nc * 4, nc, 1, 1, 1)

    def forward(self, x):
        o1 = self.act(self.c1(x))
        parts, run = [], None
        for rate in range(1, 9):
            d = getattr(self, "d%d" % rate)(o1)
            run = d if run is None else run + d
            parts.append(run)
        return x + self.c2(self.act(torch.cat(parts, 1))).mul(0.2)


class RRBlock_32(nn.Module):
    def __init__(self):
        super().__init__()
        self.RB1 = _ResBlock_32()
        self.RB2 = _ResBlock_32()
        self.RB3 = _ResBlock_32()

    def forward(self, x):
        return self.RB3(self.RB2(self.RB1(x))).mul(0.2) + x


class PPON(NativeEngineMixin, nn.Module):
    _engine_class = "PPONEngine"

    def __init__(self, in_nc, nf, nb, out_nc, upscale=4, act_type="lrelu", alpha=1.0):
        super().__init__()
        if nf != 64:
            raise NotImplementedError("PPON: RRBlock_32 is hard-wired to 64 channels, nf must be 64")
        self.alpha = alpha
        n_upscale = 1 if upscale == 3 else int(math.log(upscale, 2))
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, nb=nb, scale=upscale, alpha=float(alpha))
        # construction order = the reference's (the default initialisation consumes the RNG in this order)
        fea = B.conv_layer(in_nc, nf, kernel_size=3)
        blocks = [RRBlock_32() for _ in range(nb)]
        lr_conv = B.conv_layer(nf, nf, kernel_size=3)
        ssim = [RRBlock_32() for _ in range(2)]
        gan = [RRBlock_32() for _ in range(2)]
        f = 3 if upscale == 3 else 2
        ups = [[B.upconv_block(nf, nf, f, act_type=act_type) for _ in range(n_upscale)] for _ in range(3)]
        heads = [(B.conv_block(nf, nf, kernel_size=3, norm_type=None, act_type=act_type),
                  B.conv_block(nf, out_nc, kernel_size=3, norm_type=None, act_type=None)) for _ in range(3)]
        self.CFEM = B.sequential(fea, B.ShortcutBlock(B.sequential(*blocks, lr_conv)))
        self.SFEM = B.sequential(*ssim)
        self.PFEM = B.sequential(*gan)
        self.CRM = B.sequential(*ups[0], *heads[0])
        self.SRM = B.sequential(*ups[1], *heads[1])
        self.PRM = B.sequential(*ups[2], *heads[2])

    def forward(self, x):
        if x.is_cuda:
            return None, None, self._engine(x.device, x.dtype).forward(x)
        out_cfem = self.CFEM(x)
        out_c = self.CRM(out_cfem)
        out_sfem = self.SFEM(out_cfem)
        out_s = self.SRM(out_sfem) + out_c
        out_p = self.alpha * self.PRM(self.PFEM(out_sfem)) + out_s
        return out_c, out_s, out_p
