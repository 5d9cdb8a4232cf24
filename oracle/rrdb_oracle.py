"""CPU oracle for the RRDB/ESRGAN hot path of victorca25/iNNfer.

THIS IS TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import it; the product (innfer_b200/) never does.

It is a functional restatement (plain torch fp32 ops + numpy) of the reference algorithm; every
function cites the reference file:line it follows (paths relative to the reference tree).
Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, imported unmodified from /root/reference in the
authoring container by tools/make_golden.py; the resulting fixtures live in tests/golden/ and
tests/test_oracle_golden.py checks this file against them on every CPU run.  The OpenCV pieces of
color_fix (cv2.resize INTER_CUBIC, cv2.GaussianBlur) are third-party code absent from the
reference tree (the reference pins no version; 4.13.0 is installed here): they are restated from
OpenCV's published algorithm and pinned the same way.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- synthetic weights


def _conv_init(cout, cin, k=3):
    """Same RNG consumption as nn.Conv2d.__init__ (kaiming_uniform_(a=sqrt(5)) then bias)."""
    conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2)
    return conv.weight.detach().clone(), conv.bias.detach().clone()


def upconv_indices(scale):
    """Flat nn.Sequential indices of the tail convs (block.py:197-210 flattening)."""
    n_up = {1: 0, 2: 1, 3: 1, 4: 2, 8: 3}[scale]
    ups = [3 + 3 * i for i in range(n_up)]
    return ups, 2 + 3 * n_up, 4 + 3 * n_up


def make_state_dict(scale=4, nb=23, nf=64, in_nc=3, out_nc=3, seed=0, last_bias=0.5, plus=False):
    """Synthetic weights recipe of SURVEY.md 8(d): torch default init in the construction order of
    RRDBNet.__init__ (RRDBNet_arch.py:25-48), then the last conv's bias set to 0.5 so the uint8
    comparison is not vacuous."""
    torch.manual_seed(seed)
    sd = OrderedDict()

    def put(name, cout, cin):
        w, b = _conv_init(cout, cin)
        sd[name + ".weight"] = w
        sd[name + ".bias"] = b

    put("model.0", nf, in_nc)
    for b in range(nb):
        for r in (1, 2, 3):
            if plus:  # ESRGAN+: conv1x1 is constructed before conv1..5 (RRDBNet_arch.py:127-133), no bias
                sd["model.1.sub.%d.RDB%d.conv1x1.weight" % (b, r)] = \
                    torch.nn.Conv2d(nf, 32, 1, bias=False).weight.detach().clone()
            for k in range(5):
                put("model.1.sub.%d.RDB%d.conv%d.0" % (b, r, k + 1), 32 if k < 4 else nf, nf + 32 * k)
    put("model.1.sub.%d" % nb, nf, nf)
    ups, hr0, hr1 = upconv_indices(scale)
    for i in ups:
        put("model.%d" % i, nf, nf)
    put("model.%d" % hr0, nf, nf)
    put("model.%d" % hr1, out_nc, nf)
    if last_bias is not None:
        sd["model.%d.bias" % hr1].fill_(last_bias)
    return sd


def infer_params(sd):
    """nb / nf / in_nc / out_nc / scale / plus from the key names (run.py:103-149)."""
    scale2x, n_uplayer, out_nc, nb, plus = 0, 0, None, None, False
    for key in sd:
        parts = key.split(".")
        if len(parts) == 5 and parts[2] == "sub":
            nb = int(parts[3])
        elif len(parts) == 3:
            num = int(parts[1])
            if num > 6 and parts[0] == "model" and parts[2] == "weight":
                scale2x += 1
            if num > n_uplayer:
                n_uplayer = num
                out_nc = sd[key].shape[0]
        if "conv1x1" in key:
            plus = True
    return dict(nb=nb, nf=sd["model.0.weight"].shape[0], in_nc=sd["model.0.weight"].shape[1],
                out_nc=out_nc, scale=2 ** scale2x, plus=plus)


# ----------------------------------------------------------------------------- network forward


def _conv(sd, name, x, act=False):
    """conv_block with pad_type='zero', norm None, mode 'CNA' (block.py:213-254): Conv2d(3,1,1)
    [+ LeakyReLU(0.2)] (block.py:88-90)."""
    y = F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=1)
    return F.leaky_relu(y, 0.2) if act else y


def rdb_forward(sd, prefix, x):
    """ResidualDenseBlock_5C.forward (RRDBNet_arch.py:152-165); the ESRGAN+ additions (155-160) are
    taken when the block has a conv1x1 weight."""
    plus = (prefix + ".conv1x1.weight") in sd
    x1 = _conv(sd, prefix + ".conv1.0", x, True)
    x2 = _conv(sd, prefix + ".conv2.0", torch.cat((x, x1), 1), True)
    if plus:
        x2 = x2 + F.conv2d(x, sd[prefix + ".conv1x1.weight"])
    x3 = _conv(sd, prefix + ".conv3.0", torch.cat((x, x1, x2), 1), True)
    x4 = _conv(sd, prefix + ".conv4.0", torch.cat((x, x1, x2, x3), 1), True)
    if plus:
        x4 = x4 + x2
    x5 = _conv(sd, prefix + ".conv5.0", torch.cat((x, x1, x2, x3, x4), 1), False)
    return x5 * 0.2 + x


def rrdb_forward(sd, prefix, x):
    """RRDB.forward (RRDBNet_arch.py:91-98)."""
    out = rdb_forward(sd, prefix + ".RDB1", x)
    out = rdb_forward(sd, prefix + ".RDB2", out)
    out = rdb_forward(sd, prefix + ".RDB3", out)
    return out * 0.2 + x


def rrdbnet_forward(sd, x, scale=None):
    """RRDBNet.forward with outm=None (RRDBNet_arch.py:25-62): fea_conv, ShortcutBlock(RRDB x nb,
    LR_conv) (block.py:183-191), upconv blocks = nearest Upsample + conv + LeakyReLU
    (block.py:348-361), HR_conv0 + LeakyReLU, HR_conv1."""
    p = infer_params(sd)
    if scale is None:
        scale = p["scale"]
    nb = p["nb"]
    with torch.no_grad():
        fea = _conv(sd, "model.0", x)
        t = fea
        for b in range(nb):
            t = rrdb_forward(sd, "model.1.sub.%d" % b, t)
        t = fea + _conv(sd, "model.1.sub.%d" % nb, t)
        ups, hr0, hr1 = upconv_indices(scale)
        f = 3 if scale == 3 else 2
        for i in ups:
            t = F.interpolate(t, scale_factor=float(f), mode="nearest")
            t = _conv(sd, "model.%d" % i, t, True)
        t = _conv(sd, "model.%d" % hr0, t, True)
        return _conv(sd, "model.%d" % hr1, t, False)


# ----------------------------------------------------------------------------- SRResNet (8f rank 1)


def make_srresnet_state_dict(scale=4, nb=16, nf=64, in_nc=3, out_nc=3, seed=0, last_bias=0.5):
    """Default-initialised SRResNet weights in the construction order of SRResNet.__init__
    (SRResNet_arch.py:24-45) with the pixelshuffle upsampler of utils/defaults.py:64."""
    torch.manual_seed(seed)
    sd = OrderedDict()

    def put(name, cout, cin):
        w, b = _conv_init(cout, cin)
        sd[name + ".weight"] = w
        sd[name + ".bias"] = b

    put("model.0", nf, in_nc)
    for b in range(nb):
        put("model.1.sub.%d.res.0" % b, nf, nf)
        put("model.1.sub.%d.res.2" % b, nf, nf)
    put("model.1.sub.%d" % nb, nf, nf)
    n_up = {1: 0, 2: 1, 3: 1, 4: 2, 8: 3}[scale]
    r = 3 if scale == 3 else 2
    for i in range(n_up):
        put("model.%d" % (2 + 3 * i), nf * r * r, nf)
    put("model.%d" % (2 + 3 * n_up), nf, nf)
    put("model.%d" % (4 + 3 * n_up), out_nc, nf)
    if last_bias is not None:
        sd["model.%d.bias" % (4 + 3 * n_up)].fill_(last_bias)
    return sd


def srresnet_forward(sd, x, scale, res_scale=1.0):
    """SRResNet.forward (SRResNet_arch.py:47-62) for norm None / ReLU / CNA / pixelshuffle:
    ResNetBlock = x + res_scale * conv1(relu(conv0(x))) (77-91); pixelshuffle_block = conv,
    PixelShuffle, ReLU (block.py:333-346)."""
    def conv(name, t, relu=False):
        y = F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], padding=1)
        return F.relu(y) if relu else y
    nb = max(int(k.split(".")[3]) for k in sd if k.startswith("model.1.sub.") and k.count(".") == 4)
    n_up = {1: 0, 2: 1, 3: 1, 4: 2, 8: 3}[scale]
    r = 3 if scale == 3 else 2
    with torch.no_grad():
        fea = conv("model.0", x)
        t = fea
        for b in range(nb):
            t = t + res_scale * conv("model.1.sub.%d.res.2" % b, conv("model.1.sub.%d.res.0" % b, t, True))
        t = fea + conv("model.1.sub.%d" % nb, t)
        for i in range(n_up):
            t = F.relu(F.pixel_shuffle(conv("model.%d" % (2 + 3 * i), t), r))
        t = conv("model.%d" % (2 + 3 * n_up), t, True)
        return conv("model.%d" % (4 + 3 * n_up), t)


# ----------------------------------------------------------------------------- PPON


def ppon_tail_indices(scale):
    """Flat nn.Sequential indices inside CRM / SRM / PRM: upconv convs, HR_conv0, HR_conv1 (block.py:197-210)."""
    n_up = {1: 0, 2: 1, 3: 1, 4: 2, 8: 3}[scale]
    return [1 + 3 * i for i in range(n_up)], 3 * n_up, 3 * n_up + 2


def make_ppon_state_dict(scale=4, nb=24, nf=64, in_nc=3, out_nc=3, seed=0, last_bias=0.5):
    """Default-initialised PPON weights in the construction order of PPON.__init__ (PPON_arch.py:24-48) and
    _ResBlock_32.__init__ (79-92); the three HR_conv1 biases are set to last_bias / 0 / 0 so that out_p is centred."""
    torch.manual_seed(seed)
    sd = OrderedDict()

    def put(name, cout, cin, k=3):
        w, b = _conv_init(cout, cin, k)
        sd[name + ".weight"] = w
        sd[name + ".bias"] = b

    def rrblock(prefix):
        for r in (1, 2, 3):
            put("%s.RB%d.c1" % (prefix, r), nf, nf)
            for d in range(1, 9):
                put("%s.RB%d.d%d" % (prefix, r, d), nf // 2, nf)
            put("%s.RB%d.c2" % (prefix, r), nf, nf * 4, 1)

    put("CFEM.0", nf, in_nc)
    for b in range(nb):
        rrblock("CFEM.1.sub.%d" % b)
    put("CFEM.1.sub.%d" % nb, nf, nf)
    for b in range(2):
        rrblock("SFEM.%d" % b)
    for b in range(2):
        rrblock("PFEM.%d" % b)
    ups, hr0, hr1 = ppon_tail_indices(scale)
    for name in ("CRM", "SRM", "PRM"):
        for i in ups:
            put("%s.%d" % (name, i), nf, nf)
    for name in ("CRM", "SRM", "PRM"):
        put("%s.%d" % (name, hr0), nf, nf)
        put("%s.%d" % (name, hr1), out_nc, nf)
    if last_bias is not None:
        sd["CRM.%d.bias" % hr1].fill_(last_bias)
    # state_dict() order of the module tree: CFEM, SFEM, PFEM, CRM, SRM, PRM
    order = [k for p in ("CFEM", "SFEM", "PFEM", "CRM", "SRM", "PRM") for k in sd if k.startswith(p)]
    return OrderedDict((k, sd[k]) for k in order)


def ppon_resblock(sd, prefix, x):
    """_ResBlock_32.forward (PPON_arch.py:94-114)."""
    def conv(name, t, dil=1, k=3):
        return F.conv2d(t, sd["%s.%s.weight" % (prefix, name)], sd["%s.%s.bias" % (prefix, name)],
                        padding=(k // 2) * dil, dilation=dil)
    o1 = F.leaky_relu(conv("c1", x), 0.2)
    parts, run = [], None
    for d in range(1, 9):
        t = conv("d%d" % d, o1, d)
        run = t if run is None else run + t
        parts.append(run)
    return x + 0.2 * conv("c2", F.leaky_relu(torch.cat(parts, 1), 0.2), 1, 1)


def ppon_rrblock(sd, prefix, x):
    """RRBlock_32.forward (PPON_arch.py:123-127)."""
    t = x
    for r in (1, 2, 3):
        t = ppon_resblock(sd, "%s.RB%d" % (prefix, r), t)
    return 0.2 * t + x


def ppon_forward(sd, x, scale, alpha=1.0):
    """PPON.forward (PPON_arch.py:64-76): returns (out_c, out_s, out_p)."""
    nb = max(int(k.split(".")[3]) for k in sd if k.startswith("CFEM.1.sub.") and k.count(".") == 4)
    ups, hr0, hr1 = ppon_tail_indices(scale)
    f = 3 if scale == 3 else 2

    def conv(name, t, act=False):
        y = F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], padding=1)
        return F.leaky_relu(y, 0.2) if act else y

    def tail(name, t):
        for i in ups:
            t = conv("%s.%d" % (name, i), F.interpolate(t, scale_factor=float(f), mode="nearest"), True)
        return conv("%s.%d" % (name, hr1), conv("%s.%d" % (name, hr0), t, True))

    with torch.no_grad():
        fea = conv("CFEM.0", x)
        t = fea
        for b in range(nb):
            t = ppon_rrblock(sd, "CFEM.1.sub.%d" % b, t)
        cfem = fea + conv("CFEM.1.sub.%d" % nb, t)
        out_c = tail("CRM", cfem)
        sfem = ppon_rrblock(sd, "SFEM.1", ppon_rrblock(sd, "SFEM.0", cfem))
        out_s = tail("SRM", sfem) + out_c
        pfem = ppon_rrblock(sd, "PFEM.1", ppon_rrblock(sd, "PFEM.0", sfem))
        out_p = alpha * tail("PRM", pfem) + out_s
    return out_c, out_s, out_p


# ----------------------------------------------------------------------------- PAN


def pan_upsample_indices(scale):
    """Indices of (upconv, PA.conv, HRconv) inside PAN.upsample and whether HRconv is followed by the LeakyReLU.

    pa_upconv_block (PAN_arch.py:11-20) puts the SAME LeakyReLU instance twice into its Sequential.  With one
    block (scale 2, 3) B.sequential returns that Sequential as is (block.py:199-202) and both run; with several
    blocks it is flattened through ``children()`` (block.py:204-207), which yields a module once, so the
    activation after HRconv disappears and a block contributes five entries."""
    n_up = {1: 0, 2: 1, 3: 1, 4: 2, 8: 3}[scale]
    if n_up == 1:
        return [(1, 2, 4)], True
    return [(5 * i + 1, 5 * i + 2, 5 * i + 4) for i in range(n_up)], False


def make_pan_state_dict(scale=4, nb=16, nf=40, unf=24, in_nc=3, out_nc=3, seed=0, gamma=0.7, self_attention=True,
                        double_scpa=False):
    """Default-initialised PAN weights in the construction order of PAN.__init__ (PAN_arch.py:107-169): conv_first,
    SCPA blocks (conv1_a, conv1_b, k1, PACnv k2/k3/k4, conv3 -- SCPA.__init__ 66-84), trunk_conv, the FSA Conv1d
    projections (block.py:421-431), the upsampler blocks, conv_last.  FSA.gamma initialises to 0 (which would switch
    the attention branch off), so it is set to ``gamma`` here."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    if scale == 1:
        unf = nf
    gw = nf // 2

    def put(name, cout, cin, k=3, bias=True, conv1d=False):
        if conv1d:
            conv = torch.nn.Conv1d(cin, cout, 1)
        else:
            conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2, bias=bias)
        sd[name + ".weight"] = conv.weight.detach().clone()
        if bias:
            sd[name + ".bias"] = conv.bias.detach().clone()

    def trunk(name):
        for b in range(nb):
            pre = "%s.%d." % (name, b)
            put(pre + "conv1_a", gw, nf, 1, False)
            put(pre + "conv1_b", gw, nf, 1, False)
            put(pre + "k1.0", gw, gw, 3, False)
            put(pre + "PACnv.k2", gw, gw, 1)
            put(pre + "PACnv.k3", gw, gw, 3, False)
            put(pre + "PACnv.k4", gw, gw, 3, False)
            put(pre + "conv3", nf, 2 * gw, 1, False)

    put("conv_first", nf, in_nc)
    trunk("SCPA_trunk")
    put("trunk_conv", nf, nf)
    if double_scpa:
        trunk("SCPA_trunk2")
        put("trunk_conv2", nf, nf)
    if self_attention:
        sd["FSA.gamma"] = torch.full((1,), float(gamma))
        put("FSA.conv_f", nf // 8, nf, conv1d=True)
        put("FSA.conv_g", nf // 8, nf, conv1d=True)
        put("FSA.conv_h", nf, nf, conv1d=True)
    blocks, _ = pan_upsample_indices(scale)
    cin = nf
    for (iu, ia, ih) in blocks:
        put("upsample.%d" % iu, unf, cin)
        put("upsample.%d.conv" % ia, unf, unf, 1)
        put("upsample.%d" % ih, unf, unf)
        cin = unf
    put("conv_last", out_nc, unf)
    return sd


def pan_self_attention(sd, x, poolsize=4):
    """SelfAttentionBlock.forward with max_pool (block.py:434-473): attention over the 4x4-max-pooled map, bicubic
    resize of the result back to the input size, gamma * out + input."""
    b, c, hgt, wid = x.shape
    p = F.max_pool2d(x, poolsize, poolsize)
    n = p.shape[2] * p.shape[3]
    v = p.reshape(b, c, n)
    f = torch.einsum("oc,bcn->bon", sd["FSA.conv_f.weight"][:, :, 0], v) + sd["FSA.conv_f.bias"][None, :, None]
    g = torch.einsum("oc,bcn->bon", sd["FSA.conv_g.weight"][:, :, 0], v) + sd["FSA.conv_g.bias"][None, :, None]
    h = torch.einsum("oc,bcn->bon", sd["FSA.conv_h.weight"][:, :, 0], v) + sd["FSA.conv_h.bias"][None, :, None]
    s = torch.einsum("bci,bcj->bij", f, g)            # s[i, j] = <f_i, g_j>
    att = torch.softmax(s, dim=-1)
    o = torch.einsum("bcj,bij->bci", h, att).reshape(b, c, p.shape[2], p.shape[3])
    o = F.interpolate(o, size=(hgt, wid), mode="bicubic", align_corners=False)
    return sd["FSA.gamma"] * o + x


def pan_scpa(sd, pre, x):
    """SCPA.forward (PAN_arch.py:86-103) with PACnv.forward (48-57)."""
    def c(name, t, k):
        return F.conv2d(t, sd[pre + name + ".weight"], sd.get(pre + name + ".bias"), padding=k // 2)

    a = F.leaky_relu(c("conv1_a", x, 1), 0.2)
    bb = F.leaky_relu(c("conv1_b", x, 1), 0.2)
    a = F.leaky_relu(c("k1.0", a, 3), 0.2)
    y = torch.sigmoid(c("PACnv.k2", bb, 1))
    bb = F.leaky_relu(c("PACnv.k4", c("PACnv.k3", bb, 3) * y, 3), 0.2)
    return c("conv3", torch.cat([a, bb], 1), 1) + x


def pan_forward(sd, x, scale):
    """PAN.forward (PAN_arch.py:171-222), 'nearest' upsampler."""
    def conv(name, t):
        w = sd[name + ".weight"]
        return F.conv2d(t, w, sd.get(name + ".bias"), padding=w.shape[-1] // 2)

    def trunk(name, conv_name, t):
        nb = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith(name + "."))
        for b in range(nb):
            t = pan_scpa(sd, "%s.%d." % (name, b), t)
        return conv(conv_name, t)

    with torch.no_grad():
        fea = conv("conv_first", x)
        t = trunk("SCPA_trunk", "trunk_conv", fea)
        if "trunk_conv2.weight" in sd:
            t = trunk("SCPA_trunk2", "trunk_conv2", t)
        fea = fea + t
        if "FSA.gamma" in sd:
            fea = pan_self_attention(sd, fea)
        blocks, hr_act = pan_upsample_indices(scale)
        f = 3 if scale == 3 else 2
        for (iu, ia, ih) in blocks:
            u = conv("upsample.%d" % iu, F.interpolate(fea, scale_factor=float(f), mode="nearest"))
            u = F.leaky_relu(u * torch.sigmoid(conv("upsample.%d.conv" % ia, u)), 0.2)
            fea = conv("upsample.%d" % ih, u)
            if hr_act:
                fea = F.leaky_relu(fea, 0.2)
        out = conv("conv_last", fea)
        ilr = F.interpolate(x, scale_factor=float(scale), mode="bilinear", align_corners=True) if scale > 1 else x
        return out + ilr


# ----------------------------------------------------------------------------- tiling / blending


def tile_origins(length, p, step=0.5):
    """Window starts of tensor.unfold(dim, p, int(p*step)) plus the edge-anchored extra window
    (utils.py:349-362)."""
    s = int(p * step)
    starts = [i * s for i in range((length - p) // s + 1)]
    if (length - p) % s != 0:
        starts.append(length - p)
    return starts


def extract_patches(img, p, step=0.5):
    """extract_patches_2d(..., batch_first=True).squeeze(0) for a [1,C,H,W] image: row-major list
    of [C,p,p] crops (utils.py:349-368)."""
    ys = tile_origins(img.shape[2], p, step)
    xs = tile_origins(img.shape[3], p, step)
    return torch.stack([img[0, :, y:y + p, x:x + p] for y in ys for x in xs], 0), ys, xs


def blend_profile(P, overlap, dtype=torch.float32):
    """cat[linspace(0.1,1,overlap), ones(P-2*overlap), linspace(1,0.1,overlap)] (utils.py:413-416)."""
    return torch.cat([torch.linspace(0.1, 1.0, overlap, dtype=dtype),
                      torch.ones(P - 2 * overlap, dtype=dtype),
                      torch.linspace(1.0, 0.1, overlap, dtype=dtype)], 0)


def recompose(patches, height, width, step=0.5, scale=1):
    """recompose_tensor (utils.py:372-445) for one image."""
    full_h, full_w = scale * height, scale * width
    n, ch, P, _ = patches.shape
    overlap = scale * int(round((1.0 - step) * (P / scale)))
    eff = int(step * P)
    prof = blend_profile(P, overlap, patches.dtype)
    wpatch = prof[None, :].repeat(P, 1) * prof[:, None].repeat(1, P)
    step_int = int(P * step)
    nrow = 1 + (max(full_h, P) - P) // step_int + (1 if (max(full_h, P) - P) % step_int else 0)
    ncol = 1 + (max(full_w, P) - P) // step_int + (1 if (max(full_w, P) - P) % step_int else 0)
    wsum = torch.zeros(1, ch, full_h, full_w, dtype=patches.dtype)
    out = torch.zeros(1, ch, full_h, full_w, dtype=patches.dtype)
    idx = 0
    for h in range(nrow):
        for w in range(ncol):
            y0 = min(h * eff, full_h - P)
            x0 = min(w * eff, full_w - P)
            wsum[0, :, y0:y0 + P, x0:x0 + P] += wpatch[None]
            out[0, :, y0:y0 + P, x0:x0 + P] += patches[idx] * wpatch
            idx += 1
    return out / wsum


def chop_forward(sd, x, patch_size=200, step=0.5, forward=None, scale=None):
    """Model.chop_forward (run.py:167-202): per-tile forward with batch 1, then recompose."""
    p = min(x.shape[2], x.shape[3], patch_size)
    patches, _, _ = extract_patches(x, p, step)
    if scale is None:
        scale = infer_params(sd)["scale"]
    fwd = forward or (lambda t: rrdbnet_forward(sd, t, scale))
    outs = [fwd(patches[i:i + 1]) for i in range(patches.shape[0])]
    return recompose(torch.cat(outs, 0), x.shape[2], x.shape[3], step=step, scale=scale)


# ----------------------------------------------------------------------------- image <-> tensor


def np2tensor(img):
    """np2tensor with defaults (utils.py:164-194): uint8 HWC BGR -> float32 [1,3,H,W] RGB in [0,1]."""
    t = torch.from_numpy(np.ascontiguousarray(np.transpose(img.astype(np.float32) / 255, (2, 0, 1)))).float()
    return t.flip(-3).unsqueeze(0)


def tensor2np(t):
    """tensor2np with defaults (utils.py:197-248): [1,3,H,W] RGB -> uint8 HWC BGR,
    clip(255*x, 0, 255).round() (np.round: half to even)."""
    a = t.float().cpu().squeeze(0).flip(-3).numpy()
    a = np.transpose(a, (1, 2, 0))
    return np.clip(255 * a, 0, 255).round().astype(np.uint8)


# ----------------------------------------------------------------------------- colour fix


def srgb2linear(srgb):
    """utils/colors.py:29-46."""
    linear = np.float32(srgb) / 255.0
    return np.where(linear <= 0.04045, linear / 12.92, np.power((linear + 0.055) / 1.055, 2.4))


def linear2srgb(linear):
    """utils/colors.py:49-60 (note the truncating uint8 cast)."""
    s = np.clip(linear.copy(), 0.0, 1.0)
    s = np.where(s <= 0.0031308, s * 12.92, 1.055 * np.power(s, 1.0 / 2.4) - 0.055)
    return np.clip(s * 255.0, 0.0, 255).astype(np.uint8)


def _cubic_coeffs(x):
    """cv::interpolateCubic, A = -0.75 (OpenCV modules/imgproc/src/resize.cpp)."""
    A = np.float32(-0.75)
    x = x.astype(np.float32)
    one = np.float32(1)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def _cubic_axis(n_src, n_dst):
    scale = 1.0 / (float(n_dst) / float(n_src))
    f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    frac = (f - s.astype(np.float32)).astype(np.float32)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, n_src - 1)
    return idx, _cubic_coeffs(frac)


def resize_cubic(img, dsize):
    """cv2.resize(img, dsize=(w, h), interpolation=cv2.INTER_CUBIC) for float32 HWC: separable
    4-tap cubic convolution, half-pixel centres, replicated border, NO antialiasing when shrinking
    (OpenCV resize.cpp: HResizeCubic then VResizeCubic)."""
    w, h = dsize
    img = img.astype(np.float32)
    xi, xc = _cubic_axis(img.shape[1], w)
    yi, yc = _cubic_axis(img.shape[0], h)
    rows = (img[:, xi[:, 0]] * xc[None, :, 0, None] + img[:, xi[:, 1]] * xc[None, :, 1, None] +
            img[:, xi[:, 2]] * xc[None, :, 2, None] + img[:, xi[:, 3]] * xc[None, :, 3, None]).astype(np.float32)
    out = (rows[yi[:, 0]] * yc[:, 0, None, None] + rows[yi[:, 1]] * yc[:, 1, None, None] +
           rows[yi[:, 2]] * yc[:, 2, None, None] + rows[yi[:, 3]] * yc[:, 3, None, None])
    return out.astype(np.float32)


def gaussian_blur3(img):
    """cv2.GaussianBlur(img, (3,3), 0): separable [0.25, 0.5, 0.25], BORDER_REFLECT_101."""
    img = img.astype(np.float32)
    p = np.pad(img, ((0, 0), (1, 1), (0, 0)), mode="reflect")
    r = p[:, 1:-1] * np.float32(0.5) + (p[:, :-2] + p[:, 2:]) * np.float32(0.25)
    p = np.pad(r, ((1, 1), (0, 0), (0, 0)), mode="reflect")
    return (p[1:-1] * np.float32(0.5) + (p[:-2] + p[2:]) * np.float32(0.25)).astype(np.float32)


def color_fix(img_a, img_b):
    """color_fix (utils.py:278-315): add the low-frequency difference LR - SR(down) back to SR."""
    a = srgb2linear(img_a)
    b = srgb2linear(img_b)
    ha, wa = a.shape[:2]
    hb, wb = b.shape[:2]
    scaling = ha < hb and wa < wb
    b_ds = resize_cubic(b, (wa, ha)) if scaling else b
    blurred = gaussian_blur3(a - b_ds)
    if scaling:
        blurred = resize_cubic(blurred, (wb, hb))
    return linear2srgb(blurred + b)


# ----------------------------------------------------------------------------- key mapping


def mod2normal(sd):
    """'new-arch' ESRGAN keys -> original keys (utils.py:666-698); hard-codes 23 blocks / 4x."""
    if "conv_first.weight" not in sd:
        return sd
    out = OrderedDict()
    out["model.0.weight"] = sd["conv_first.weight"]
    out["model.0.bias"] = sd["conv_first.bias"]
    for k, v in sd.items():
        if "RDB" in k:
            nk = k.replace("RRDB_trunk.", "model.1.sub.")
            if ".weight" in k:
                nk = nk.replace(".weight", ".0.weight")
            elif ".bias" in k:
                nk = nk.replace(".bias", ".0.bias")
            out[nk] = v
    for new, old in (("model.1.sub.23", "trunk_conv"), ("model.3", "upconv1"), ("model.6", "upconv2"),
                     ("model.8", "HRconv"), ("model.10", "conv_last")):
        out[new + ".weight"] = sd[old + ".weight"]
        out[new + ".bias"] = sd[old + ".bias"]
    return out


def swa2normal(sd):
    """SWA wrapper removal (utils.py:701-720)."""
    if "n_averaged" not in sd:
        return sd
    return OrderedDict((k.replace("module.module.", ""), v) for k, v in sd.items()
                       if "n_averaged" not in k and "module.module." in k)


# ----------------------------------------------------------------------------- work accounting


def flop_per_lr_pixel(scale=4, nb=23, nf=64, in_nc=3, out_nc=3):
    """Algorithmic conv FLOPs per low-res pixel (real channels only; SURVEY.md 8d: 35 853 696 for
    the 4x net)."""
    mac = 9 * in_nc * nf
    rdb = 9 * (sum((nf + 32 * k) * 32 for k in range(4)) + (nf + 128) * nf)
    mac += nb * 3 * rdb + 9 * nf * nf
    ups, _, _ = upconv_indices(scale)
    f = 3 if scale == 3 else 2
    res = 1
    for _ in ups:
        res *= f * f
        mac += res * 9 * nf * nf
    mac += res * 9 * nf * nf + res * 9 * nf * out_nc
    return 2 * mac
