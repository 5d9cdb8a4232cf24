"""UnetGenerator (pix2pix generator) -- mirror of the reference's architectures/UNet_arch.py:11-165.

The module tree owns the parameters under the reference's state-dict names ('model.model.0.weight',
'model.model.1.model.1.weight', ... : every level is a block whose ``model`` Sequential holds
[LeakyReLU, down conv, norm, inner block, ReLU, transposed conv, norm]) and serves the explicit ``-cpu`` mode;
``forward`` on a CUDA tensor runs the sm_100a engine (csrc/i2i.cu).  Levels are created innermost first, down conv
before transposed conv, like the reference, so that a seeded construction draws the same initial weights.
"""
import torch
import torch.nn as nn

from ._native import NativeEngineMixin


def _norm_class(norm_type):
    if norm_type in ("BN", "batch"):
        return nn.BatchNorm2d
    if norm_type in ("IN", "instance"):
        return nn.InstanceNorm2d
    raise NameError("Unknown norm layer")


class UnetSkipConnectionBlock(nn.Module):
    """One level: x -> cat([x, up(inner(down(x)))]) (the outermost level returns tanh(up(..)) instead).  The
    LeakyReLU in front of the down conv is in-place as in the reference (UNet_arch.py:112), so the skip half of
    the concatenation carries leaky_relu(x)."""

    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=nn.BatchNorm2d, use_dropout=False, upsample_mode="deconv"):
        super().__init__()
        if upsample_mode != "deconv":
            raise NotImplementedError("UnetGenerator: only upsample_mode='deconv' (the reference default) is supported")
        if use_dropout:
            raise NotImplementedError("UnetGenerator: dropout is not supported (random in the pix2pix no-eval mode)")
        self.outermost = outermost
        bias = norm_layer is nn.InstanceNorm2d
        down = nn.Conv2d(outer_nc if input_nc is None else input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=bias)
        up_in = inner_nc if innermost else inner_nc * 2
        up = nn.ConvTranspose2d(up_in, outer_nc, kernel_size=4, stride=2, padding=1, bias=True if outermost else bias)
        if outermost:
            layers = [down, submodule, nn.ReLU(True), up, nn.Tanh()]
        elif innermost:
            layers = [nn.LeakyReLU(0.2, True), down, nn.ReLU(True), up, norm_layer(outer_nc)]
        else:
            layers = [nn.LeakyReLU(0.2, True), down, norm_layer(inner_nc), submodule, nn.ReLU(True), up, norm_layer(outer_nc)]
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x) if self.outermost else torch.cat([x, self.model(x)], 1)


class UnetGenerator(NativeEngineMixin, nn.Module):
    _engine_class = "UNetEngine"

    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_type="batch", use_dropout=False,
                 upsample_mode="deconv"):
        super().__init__()
        if num_downs < 5:
            raise ValueError("UnetGenerator needs num_downs >= 5")
        norm_layer = _norm_class(norm_type)
        self.cfg = dict(in_nc=input_nc, out_nc=output_nc, num_downs=num_downs, ngf=ngf,
                        norm="batch" if norm_layer is nn.BatchNorm2d else "instance", scale=1)
        kw = dict(norm_layer=norm_layer, upsample_mode=upsample_mode)
        block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, innermost=True, **kw)
        for _ in range(num_downs - 5):
            block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, submodule=block, use_dropout=use_dropout, **kw)
        for mult in (4, 2, 1):
            block = UnetSkipConnectionBlock(ngf * mult, ngf * mult * 2, submodule=block, **kw)
        self.model = UnetSkipConnectionBlock(output_nc, ngf, input_nc=input_nc, submodule=block, outermost=True, **kw)

    def _engine_key_extra(self):
        # BatchNorm uses the statistics of the batch in training mode (run.py:297: pix2pix runs with meval False)
        return (self.training,)

    def forward(self, input):
        if input.is_cuda:
            return self._engine(input.device, input.dtype).forward(input)
        return self.model(input)
