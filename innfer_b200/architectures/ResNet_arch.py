"""ResnetGenerator (CycleGAN generator) -- mirror of the reference's architectures/ResNet_arch.py:11-151.

One flat Sequential under ``model`` with the reference's indices (0 reflection pad, 1 7x7 conv, 2 norm, 3 ReLU,
4-9 the two stride-2 convs, 10.. the ResnetBlocks whose ``conv_block`` holds [pad, conv, norm, ReLU, pad, conv, norm],
then two transposed convs with norm + ReLU, pad, 7x7 conv, tanh).  ``forward`` on a CUDA tensor runs the sm_100a
engine (csrc/i2i.cu); on a CPU tensor the torch modules (explicit ``-cpu`` mode).
"""
import torch.nn as nn

from ._native import NativeEngineMixin
from .UNet_arch import _norm_class


class ResnetBlock(nn.Module):
    """x + conv_block(x) with conv_block = [pad, conv3x3, norm, ReLU, pad, conv3x3, norm]."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        if padding_type != "reflect":
            raise NotImplementedError("ResnetBlock: only padding_type='reflect' (the reference default) is supported")
        if use_dropout:
            raise NotImplementedError("ResnetBlock: dropout is not supported")
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias), norm_layer(dim), nn.ReLU(True),
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias), norm_layer(dim))

    def forward(self, x):
        return x + self.conv_block(x)


class ResnetGenerator(NativeEngineMixin, nn.Module):
    _engine_class = "ResNetGenEngine"

    def __init__(self, input_nc, output_nc, ngf=64, norm_type="batch", use_dropout=False, n_blocks=6,
                 padding_type="reflect", upsample_mode="deconv"):
        super().__init__()
        assert n_blocks >= 0
        if upsample_mode != "deconv":
            raise NotImplementedError("ResnetGenerator: only upsample_mode='deconv' (the reference default) is supported")
        norm_layer = _norm_class(norm_type)
        bias = norm_layer is nn.InstanceNorm2d
        self.cfg = dict(in_nc=input_nc, out_nc=output_nc, ngf=ngf, n_blocks=n_blocks,
                        norm="batch" if norm_layer is nn.BatchNorm2d else "instance", scale=1)
        layers = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=bias), norm_layer(ngf),
                  nn.ReLU(True)]
        ch = ngf
        for _ in range(2):
            layers += [nn.Conv2d(ch, ch * 2, kernel_size=3, stride=2, padding=1, bias=bias), norm_layer(ch * 2), nn.ReLU(True)]
            ch *= 2
        layers += [ResnetBlock(ch, padding_type, norm_layer, use_dropout, bias) for _ in range(n_blocks)]
        for _ in range(2):
            layers += [nn.ConvTranspose2d(ch, ch // 2, kernel_size=3, stride=2, padding=1, output_padding=1, bias=bias),
                       norm_layer(ch // 2), nn.ReLU(True)]
            ch //= 2
        layers += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*layers)

    def _engine_key_extra(self):
        return (self.training,)

    def forward(self, x):
        if x.is_cuda:
            return self._engine(x.device, x.dtype).forward(x)
        return self.model(x)
