"""CPU oracle of the image-to-image generators of the reference: pix2pix's UnetGenerator and
CycleGAN's ResnetGenerator.  TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product path.

Functional restatement (torch fp32 ops on the reference-named state dict) of
  architectures/UNet_arch.py:11-161   (UnetGenerator / UnetSkipConnectionBlock, upsample_mode 'deconv')
  architectures/ResNet_arch.py:11-151 (ResnetGenerator / ResnetBlock, upsample_mode 'deconv', padding 'reflect')
and of the run.py glue that differs for these families (run.py:295-361): `normalize` on (images in [-1, 1]),
pix2pix with meval=False (BatchNorm uses the statistics of the batch it is given), cyclegan with chop=True and
strict=False.  Pinned against fixtures produced by the unmodified reference (tests/golden/i2i_*.npz, written by
tools/make_golden.py i2i); torch's conv / norm arithmetic is third-party code the reference does not vendor.
"""
import torch
import torch.nn.functional as F


def unet_keys(num_downs):
    """State-dict prefixes of the levels, outermost first: level i lives under 'model' + '.model.<k>' chains.
    Sequential indices follow UNet_arch.py:120-158: outermost [down, sub, relu, up, tanh]; middle
    [lrelu, down, norm, sub, relu, up, norm]; innermost [lrelu, down, relu, up, norm]."""
    levels = []
    prefix = "model.model"
    for i in range(num_downs):
        if i == 0:
            levels.append(dict(down=prefix + ".0", up=prefix + ".3", dnorm=None, unorm=None, sub=prefix + ".1.model"))
        elif i < num_downs - 1:
            levels.append(dict(down=prefix + ".1", dnorm=prefix + ".2", up=prefix + ".5", unorm=prefix + ".6",
                               sub=prefix + ".3.model"))
        else:
            levels.append(dict(down=prefix + ".1", dnorm=None, up=prefix + ".3", unorm=prefix + ".4", sub=None))
        prefix = levels[-1]["sub"]
    return levels


def _norm(sd, name, x, kind, train):
    """nn.BatchNorm2d (affine, batch statistics when the module is in training mode -- run.py:297 meval False -- else
    the running ones) or nn.InstanceNorm2d (no affine, always per-sample statistics); eps 1e-5."""
    if kind == "instance":
        return F.instance_norm(x, eps=1e-5)
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if train:
        return F.batch_norm(x, None, None, w, b, training=True, eps=1e-5)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], w, b, training=False, eps=1e-5)


def unet_forward(sd, x, num_downs=8, norm="batch", train=True):
    """UnetGenerator.forward (UNet_arch.py:68-70,160-165).  The LeakyReLU at the head of every inner block is
    in-place (UNet_arch.py:112), so the skip connection carries leaky_relu(x), not x."""
    lv = unet_keys(num_downs)

    def bias(name):
        return sd.get(name + ".bias")

    def block(i, x):
        L = lv[i]
        if i == 0:
            d = F.conv2d(x, sd[L["down"] + ".weight"], bias(L["down"]), stride=2, padding=1)
            u = block(1, d)
            y = F.conv_transpose2d(F.relu(u), sd[L["up"] + ".weight"], bias(L["up"]), stride=2, padding=1)
            return torch.tanh(y)
        a = F.leaky_relu(x, 0.2)                       # in place in the reference: this is also the skip tensor
        d = F.conv2d(a, sd[L["down"] + ".weight"], bias(L["down"]), stride=2, padding=1)
        if L["dnorm"]:
            d = _norm(sd, L["dnorm"], d, norm, train)
        inner = block(i + 1, d) if L["sub"] else d
        y = F.conv_transpose2d(F.relu(inner), sd[L["up"] + ".weight"], bias(L["up"]), stride=2, padding=1)
        y = _norm(sd, L["unorm"], y, norm, train)
        return torch.cat([a, y], 1)

    return block(0, x)


def resnet_forward(sd, x, n_blocks=9, norm="instance", train=False):
    """ResnetGenerator.forward (ResNet_arch.py:55-93) with reflect padding and ConvTranspose2d upsampling.
    Sequential indices: 0 pad, 1 conv7, 2 norm, 3 relu; 4/5/6 and 7/8/9 the two stride-2 convs; 10 .. 10+n-1 the
    blocks (conv_block.{1,2} conv+norm, relu, .{5,6} conv+norm); then convT/norm/relu twice, pad, conv7, tanh."""
    def conv(name, t, **kw):
        return F.conv2d(t, sd[name + ".weight"], sd.get(name + ".bias"), **kw)

    def nrm(name, t):
        return _norm(sd, name, t, norm, train)

    y = F.relu(nrm("model.2", conv("model.1", F.pad(x, (3, 3, 3, 3), mode="reflect"))))
    y = F.relu(nrm("model.5", conv("model.4", y, stride=2, padding=1)))
    y = F.relu(nrm("model.8", conv("model.7", y, stride=2, padding=1)))
    for b in range(n_blocks):
        p = "model.%d.conv_block" % (10 + b)
        t = F.relu(nrm(p + ".2", conv(p + ".1", F.pad(y, (1, 1, 1, 1), mode="reflect"))))
        t = nrm(p + ".6", conv(p + ".5", F.pad(t, (1, 1, 1, 1), mode="reflect")))
        y = y + t
    i = 10 + n_blocks
    for _ in range(2):
        y = F.conv_transpose2d(y, sd["model.%d.weight" % i], sd.get("model.%d.bias" % i), stride=2, padding=1,
                               output_padding=1)
        y = F.relu(nrm("model.%d" % (i + 1), y))
        i += 3
    y = conv("model.%d" % (i + 1), F.pad(y, (3, 3, 3, 3), mode="reflect"))
    return torch.tanh(y)


def norm_range(x):
    """utils.norm (utils.py:152-161): [0, 1] -> [-1, 1], clamped."""
    return ((x - 0.5) * 2.0).clamp(-1, 1)


def denorm_range(x):
    """utils.denorm (utils.py:136-150): [-1, 1] -> [0, 1], clamped."""
    return ((x + 1.0) / 2.0).clamp(0, 1)


# ----------------------------------------------------------------------------- seeded weights
def make_unet_state_dict(num_downs=8, ngf=64, in_nc=3, out_nc=3, norm="batch", seed=0):
    """The state dict UnetGenerator(...) has after torch.manual_seed(seed): torch's default initialisation of the
    convs drawn in the reference's construction order (innermost level first, down conv before transposed conv,
    UNet_arch.py:47-66,108-116).  BatchNorm: weight 1, bias 0, running stats 0 / 1."""
    import torch.nn as nn
    torch.manual_seed(seed)
    bias = norm == "instance"
    chans = [(ngf * 8, ngf * 8)] * (num_downs - 4) + [(ngf * 4, ngf * 8), (ngf * 2, ngf * 4), (ngf, ngf * 2), (out_nc, ngf)]
    built = []                        # innermost first: (outer_nc, inner_nc, down, up)
    for idx, (outer, inner) in enumerate(chans):
        innermost, outermost = idx == 0, idx == len(chans) - 1
        down = nn.Conv2d(in_nc if outermost else outer, inner, 4, 2, 1, bias=bias)
        up = nn.ConvTranspose2d(inner if innermost else inner * 2, outer, 4, 2, 1, bias=True if outermost else bias)
        built.append((outer, inner, down, up))
    sd = {}
    levels = unet_keys(num_downs)
    for depth, L in enumerate(levels):
        outer, inner, down, up = built[len(built) - 1 - depth]
        for name, mod in ((L["down"], down), (L["up"], up)):
            sd[name + ".weight"] = mod.weight.detach().clone()
            if mod.bias is not None:
                sd[name + ".bias"] = mod.bias.detach().clone()
        if norm == "batch":
            for name, c in ((L["dnorm"], inner), (L["unorm"], outer)):
                if name:
                    sd[name + ".weight"], sd[name + ".bias"] = torch.ones(c), torch.zeros(c)
                    sd[name + ".running_mean"], sd[name + ".running_var"] = torch.zeros(c), torch.ones(c)
                    sd[name + ".num_batches_tracked"] = torch.tensor(0)
    return sd


def make_resnet_state_dict(n_blocks=9, ngf=64, in_nc=3, out_nc=3, norm="instance", seed=0):
    """Same for ResnetGenerator (ResNet_arch.py:55-91): modules are created front to back."""
    import torch.nn as nn
    torch.manual_seed(seed)
    bias = norm == "instance"
    sd = {}

    def put(name, mod):
        sd[name + ".weight"] = mod.weight.detach().clone()
        if mod.bias is not None:
            sd[name + ".bias"] = mod.bias.detach().clone()

    def put_norm(name, c):
        if norm == "batch":
            sd[name + ".weight"], sd[name + ".bias"] = torch.ones(c), torch.zeros(c)
            sd[name + ".running_mean"], sd[name + ".running_var"] = torch.zeros(c), torch.ones(c)
            sd[name + ".num_batches_tracked"] = torch.tensor(0)

    put("model.1", nn.Conv2d(in_nc, ngf, 7, bias=bias))
    put_norm("model.2", ngf)
    ch = ngf
    for i in (4, 7):
        put("model.%d" % i, nn.Conv2d(ch, ch * 2, 3, 2, 1, bias=bias))
        put_norm("model.%d" % (i + 1), ch * 2)
        ch *= 2
    for b in range(n_blocks):
        p = "model.%d.conv_block" % (10 + b)
        for j in (1, 5):
            put("%s.%d" % (p, j), nn.Conv2d(ch, ch, 3, bias=bias))
            put_norm("%s.%d" % (p, j + 1), ch)
    i = 10 + n_blocks
    for _ in range(2):
        put("model.%d" % i, nn.ConvTranspose2d(ch, ch // 2, 3, 2, 1, output_padding=1, bias=bias))
        put_norm("model.%d" % (i + 1), ch // 2)
        ch //= 2
        i += 3
    put("model.%d" % (i + 1), nn.Conv2d(ngf, out_nc, 7))
    return sd


def randomize_norms(sd, seed):
    """Make the norm layers visible to a parity test: affine weights U(0.5, 1.5), biases U(-0.3, 0.3), running means
    N(0, 0.2), running variances U(0.5, 1.5), drawn in SORTED key order from a private generator.  In place; returns sd."""
    g = torch.Generator().manual_seed(seed)
    for key in sorted(sd):
        if not key.endswith(".running_mean"):
            continue
        base = key[:-len(".running_mean")]
        c = sd[key].numel()
        sd[base + ".weight"].copy_(torch.rand(c, generator=g) + 0.5)
        sd[base + ".bias"].copy_(torch.rand(c, generator=g) * 0.6 - 0.3)
        sd[base + ".running_mean"].copy_(torch.randn(c, generator=g) * 0.2)
        sd[base + ".running_var"].copy_(torch.rand(c, generator=g) + 0.5)
    return sd
