"""SRResNet (SRGAN generator) -- mirror of the reference's architectures/SRResNet_arch.py:15-91
(SRResNet 15-62, ResNetBlock 64-91) for the configuration get_network_G_config produces
(utils/defaults.py:53-67): no norm layers, ReLU, mode CNA, pixelshuffle (or upconv) upsampling.

As for RRDBNet the module tree owns the parameters under the reference's key names and serves the
explicit ``-cpu`` mode; a CUDA tensor goes to the sm_100a engine.
"""
import math

import torch
import torch.nn as nn

from . import block as B
from ._native import NativeEngineMixin


class ResNetBlock(nn.Module):
    """x + res_scale * conv1(act(conv0(x))) (3-3 style, EDSR residual scaling)."""

    def __init__(self, in_nc, mid_nc, out_nc, kernel_size=3, stride=1, dilation=1, groups=1, bias=True,
                 pad_type="zero", norm_type=None, act_type="relu", mode="CNA", res_scale=1, convtype="Conv2D"):
        super().__init__()
        if mode != "CNA":
            raise NotImplementedError("ResNetBlock: only mode='CNA' is supported")
        conv0 = B.conv_block(in_nc, mid_nc, kernel_size, stride, dilation, groups, bias, pad_type, norm_type,
                             act_type, mode, convtype)
        conv1 = B.conv_block(mid_nc, out_nc, kernel_size, stride, dilation, groups, bias, pad_type, norm_type,
                             None, mode, convtype)
        self.res = B.sequential(conv0, conv1)
        self.res_scale = res_scale

    def forward(self, x):
        return x + self.res(x).mul(self.res_scale)


class SRResNet(NativeEngineMixin, nn.Module):
    _engine_class = "SRResNetEngine"

    def __init__(self, in_nc, out_nc, nf, nb, upscale=4, norm_type="batch", act_type="relu", mode="NAC",
                 res_scale=1, upsample_mode="upconv", convtype="Conv2D", finalact=None):
        super().__init__()
        if norm_type or mode != "CNA" or act_type.lower() != "relu" or finalact or convtype != "Conv2D":
            raise NotImplementedError("SRResNet: only norm_type=None, mode='CNA', relu, Conv2D, finalact=None "
                                      "(the get_network_G_config defaults) are supported")
        n_upscale = 1 if upscale == 3 else int(math.log(upscale, 2))
        self.cfg = dict(in_nc=in_nc, out_nc=out_nc, nf=nf, nb=nb, scale=upscale, res_scale=float(res_scale),
                        upsample_mode=upsample_mode)
        fea = B.conv_block(in_nc, nf, kernel_size=3, norm_type=None, act_type=None)
        blocks = [ResNetBlock(nf, nf, nf, norm_type=norm_type, act_type=act_type, mode=mode, res_scale=res_scale,
                              convtype=convtype) for _ in range(nb)]
        lr_conv = B.conv_block(nf, nf, kernel_size=3, norm_type=norm_type, act_type=None, mode=mode)
        if upsample_mode == "upconv":
            up_block = B.upconv_block
        elif upsample_mode == "pixelshuffle":
            up_block = B.pixelshuffle_block
        else:
            raise NotImplementedError("upsample mode [{:s}] is not found".format(upsample_mode))
        if upscale == 3:
            ups = [up_block(nf, nf, 3, act_type=act_type)]
        else:
            ups = [up_block(nf, nf, act_type=act_type) for _ in range(n_upscale)]
        hr0 = B.conv_block(nf, nf, kernel_size=3, norm_type=None, act_type=act_type)
        hr1 = B.conv_block(nf, out_nc, kernel_size=3, norm_type=None, act_type=None)
        self.model = B.sequential(fea, B.ShortcutBlock(B.sequential(*blocks, lr_conv)), *ups, hr0, hr1)

    def forward(self, x, outm=None):
        y = self._engine(x.device, x.dtype).forward(x) if x.is_cuda else self.model(x)
        if outm == "scaltanh":
            return (torch.tanh(y) + 1.0) / 2.0
        if outm == "tanh":
            return torch.tanh(y)
        if outm == "sigmoid":
            return torch.sigmoid(y)
        if outm == "clamp":
            return torch.clamp(y, min=0.0, max=1.0)
        return y
