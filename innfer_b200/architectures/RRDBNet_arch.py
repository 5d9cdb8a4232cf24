"""RRDBNet (ESRGAN generator, original key layout) -- mirror of the reference's
architectures/RRDBNet_arch.py:16-165 (RRDBNet 16-62, RRDB 64-98, ResidualDenseBlock_5C 100-165).

The module tree only exists to own the parameters under the reference's state-dict names and to
serve the explicit ``-cpu`` mode.  ``forward`` on a CUDA tensor runs the sm_100a engine.
"""
import math

import torch
import torch.nn as nn

from . import block as B
from ._native import NativeEngineMixin


class ResidualDenseBlock_5C(nn.Module):
    def __init__(self, nf=64, kernel_size=3, gc=32, stride=1, bias=1, pad_type="zero", norm_type=None,
                 act_type="leakyrelu", mode="CNA", convtype="Conv2D", spectral_norm=False,
                 gaussian_noise=False, plus=False):
        super().__init__()
        self.noise = B.GaussianNoise() if gaussian_noise else None
        self.conv1x1 = B.conv1x1(nf, gc) if plus else None
        kw = dict(bias=bias, pad_type=pad_type, norm_type=norm_type, mode=mode, convtype=convtype,
                  spectral_norm=spectral_norm)
        self.conv1 = B.conv_block(nf, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv2 = B.conv_block(nf + gc, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv3 = B.conv_block(nf + 2 * gc, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv4 = B.conv_block(nf + 3 * gc, gc, kernel_size, stride, act_type=act_type, **kw)
        self.conv5 = B.conv_block(nf + 4 * gc, nf, 3, stride, act_type=None, **kw)

    def forward(self, x):
        x1 = self.conv1(x)
        x2 = self.conv2(torch.cat((x, x1), 1))
        if self.conv1x1 is not None:
            x2 = x2 + self.conv1x1(x)
        x3 = self.conv3(torch.cat((x, x1, x2), 1))
        x4 = self.conv4(torch.cat((x, x1, x2, x3), 1))
        if self.conv1x1 is not None:
            x4 = x4 + x2
        x5 = self.conv5(torch.cat((x, x1, x2, x3, x4), 1))
        out = x5 * 0.2 + x
        return self.noise(out) if self.noise is not None else out


class RRDB(nn.Module):
    def __init__(self, nf, nr=3, kernel_size=3, gc=32, stride=1, bias=1, pad_type="zero", norm_type=None,
                 act_type="leakyrelu", mode="CNA", convtype="Conv2D", spectral_norm=False,
                 gaussian_noise=False, plus=False):
        super().__init__()
        if nr != 3:
            raise NotImplementedError("RRDB with nr != 3 is not supported (reference default nr=3)")
        mk = lambda: ResidualDenseBlock_5C(nf, kernel_size, gc, stride, bias, pad_type, norm_type, act_type,  # noqa: E731
                                           mode, convtype, spectral_norm=spectral_norm,
                                           gaussian_noise=gaussian_noise, plus=plus)
        self.RDB1, self.RDB2, self.RDB3 = mk(), mk(), mk()

    def forward(self, x):
        return self.RDB3(self.RDB2(self.RDB1(x))) * 0.2 + x


class RRDBNet(NativeEngineMixin, nn.Module):
    _engine_class = "RRDBEngine"

    def __init__(self, in_nc, out_nc, nf, nb, nr=3, gc=32, upscale=4, norm_type=None, act_type="leakyrelu",
                 mode="CNA", upsample_mode="upconv", convtype="Conv2D", finalact=None, gaussian_noise=False,
                 plus=False):
        super().__init__()
        if norm_type or act_type.lower() not in ("leakyrelu", "lrelu") or upsample_mode != "upconv" or finalact:
            raise NotImplementedError("RRDBNet: only norm_type=None, leakyrelu, upconv, finalact=None is supported")
        n_upscale = int(math.log(upscale, 2))
        if upscale == 3:
            n_upscale = 1
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, nb=nb, gc=gc, scale=upscale, plus=bool(plus))
        fea = B.conv_block(in_nc, nf, kernel_size=3, norm_type=None, act_type=None, convtype=convtype)
        # like the reference (RRDBNet_arch.py:26) the blocks always use 32 growth channels
        blocks = [RRDB(nf, nr, kernel_size=3, gc=32, stride=1, bias=1, pad_type="zero", norm_type=norm_type,
                       act_type=act_type, mode="CNA", convtype=convtype, gaussian_noise=gaussian_noise, plus=plus)
                  for _ in range(nb)]
        lr_conv = B.conv_block(nf, nf, kernel_size=3, norm_type=norm_type, act_type=None, mode=mode, convtype=convtype)
        if upscale == 3:
            ups = [B.upconv_block(nf, nf, 3, act_type=act_type, convtype=convtype)]
        else:
            ups = [B.upconv_block(nf, nf, act_type=act_type, convtype=convtype) for _ in range(n_upscale)]
        hr0 = B.conv_block(nf, nf, kernel_size=3, norm_type=None, act_type=act_type, convtype=convtype)
        hr1 = B.conv_block(nf, out_nc, kernel_size=3, norm_type=None, act_type=None, convtype=convtype)
        self.model = B.sequential(fea, B.ShortcutBlock(B.sequential(*blocks, lr_conv)), *ups, hr0, hr1)

    def forward(self, x, outm=None):
        if x.is_cuda:
            y = self._engine(x.device, x.dtype).forward(x)
        else:
            y = self.model(x)  # explicit -cpu mode
        if outm == "scaltanh":
            return (torch.tanh(y) + 1.0) / 2.0
        if outm == "tanh":
            return torch.tanh(y)
        if outm == "sigmoid":
            return torch.sigmoid(y)
        if outm == "clamp":
            return torch.clamp(y, min=0.0, max=1.0)
        return y
