"""Generates tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the authoring container only (the GPU box has no /root/reference):
    python tools/make_golden.py
Each fixture stores the seeded inputs' recipe and the reference outputs, so that
tests/test_oracle_golden.py can pin oracle/rrdb_oracle.py and the GPU tests can compare the CUDA
path with reference outputs without the mount.
"""
import os
import sys
import tempfile

import numpy as np
import torch

REF = os.environ.get("INNFER_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import run as ref_run  # noqa: E402
from architectures import get_network  # noqa: E402
from utils import utils as ref_utils  # noqa: E402
from utils.defaults import get_network_G_config  # noqa: E402

from oracle import rrdb_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def ref_net(scale, nb, nf=64, seed=0, last_bias=0.5, plus=False):
    torch.manual_seed(seed)
    cfg = get_network_G_config({"type": "esrgan", "nb": nb, "nf": nf, "plus": plus}, scale)
    net = get_network(cfg).eval()
    _, _, hr1 = O.upconv_indices(scale)
    with torch.no_grad():
        net.state_dict()["model.%d.bias" % hr1].fill_(last_bias)
    return net


def image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def save_model(net, path):
    torch.save(net.state_dict(), path)


def ppon_fixtures(td):
    # ---- G4d: PPON: auto-detected from 'CFEM.0.weight', third output only (run.py:191-192)
    for scale, nb, (h, w) in ((4, 1, (40, 48)), (2, 2, (36, 44))):
        torch.manual_seed(17)
        net = get_network(get_network_G_config({"type": "ppon", "nb": nb}, scale)).eval()
        _, _, hr1 = O.ppon_tail_indices(scale)
        with torch.no_grad():
            net.state_dict()["CRM.%d.bias" % hr1].fill_(0.5)
        path = os.path.join(td, "%dx_ppon.pth" % scale)
        save_model(net, path)
        # Model(...,'infer') keeps the default depth (nb = 24) for ppon (run.py:157-163) and would reject a shallower
        # checkpoint: wrap the network directly, chop_forward only needs these attributes
        model = ref_run.Model.__new__(ref_run.Model)
        model.arch, model.scale, model.model, model.chop, model.device = "ppon", scale, net, True, torch.device("cpu")
        img = image(18, h, w)
        y = model.chop_forward(ref_utils.np2tensor(img), patch_size=32, step=0.5)
        with torch.no_grad():
            oc, os_, op = net(ref_utils.np2tensor(img))
        np.savez_compressed(os.path.join(OUT, "ppon_s%d_nb%d_%dx%d_p32.npz" % (scale, nb, h, w)), img_seed=18,
                            h=h, w=w, patch=32, seed=17, scale=scale, nb=nb, arch=model.arch,
                            keys=np.array(list(net.state_dict().keys())),
                            wsum=np.array([float(v.double().sum()) for v in net.state_dict().values()]),
                            y=y.numpy().astype(np.float32), u8=ref_utils.tensor2np(y.detach()),
                            **({} if scale == 4 else dict(out_c=oc.numpy().astype(np.float32),
                                                          out_s=os_.numpy().astype(np.float32),
                                                          out_p=op.numpy().astype(np.float32))))


def pan_fixtures(td):
    # ---- G4e: PAN: auto-detected from 'SCPA_trunk.0.conv1_a.weight' (run.py:50-53); pixel and
    # self attention, bicubic / bilinear resampling.  Scale 2 keeps the LeakyReLU after HRconv, scale 4 loses it.
    for scale, nb, (h, w) in ((4, 2, (40, 48)), (2, 1, (36, 44)), (3, 1, (24, 28)), (1, 1, (33, 40))):
        torch.manual_seed(23)
        net = get_network(get_network_G_config({"type": "pan", "nb": nb}, scale)).eval()
        with torch.no_grad():
            net.FSA.gamma.fill_(0.7)
        model = ref_run.Model.__new__(ref_run.Model)
        model.arch, model.scale, model.model, model.chop, model.device = "pan", scale, net, True, torch.device("cpu")
        img = image(24, h, w)
        y = model.chop_forward(ref_utils.np2tensor(img), patch_size=32, step=0.5)
        with torch.no_grad():
            whole = net(ref_utils.np2tensor(img))
        np.savez_compressed(os.path.join(OUT, "pan_s%d_nb%d_%dx%d_p32.npz" % (scale, nb, h, w)), img_seed=24,
                            h=h, w=w, patch=32, seed=23, gamma=0.7, scale=scale, nb=nb, arch=model.arch,
                            keys=np.array(list(net.state_dict().keys())),
                            wsum=np.array([float(v.double().sum()) for v in net.state_dict().values()]),
                            y=y.numpy().astype(np.float32), u8=ref_utils.tensor2np(y.detach()),
                            whole=whole.numpy().astype(np.float32))


def steps_fixtures(td):
    # ---- G3b: chop_forward / recompose_tensor with step != 0.5 (Model.chop_forward's own default is 1.0,
    # run.py:167; recompose_tensor accepts [0.5, 1.0], utils.py:391): overlap, effective stride and the
    # edge-anchored extra tiles all change with the step
    out = {}
    for scale, nb, (h, w), patch, step in ((4, 1, (40, 56), 32, 1.0), (4, 1, (40, 56), 32, 0.75), (2, 1, (50, 70), 32, 0.75),
                                           (1, 1, (80, 64), 32, 0.625), (2, 1, (44, 36), 20, 0.9), (4, 1, (30, 52), 200, 1.0)):
        net = ref_net(scale, nb, seed=21)
        path = os.path.join(td, "%dx_steps.pth" % scale)
        save_model(net, path)
        model = ref_run.Model(path, "infer", None, device=torch.device("cpu"), chop=True)
        img = image(31, h, w)
        y = model.chop_forward(ref_utils.np2tensor(img), patch_size=patch, step=step)
        key = "s%d_%dx%d_p%d_step%s" % (scale, h, w, patch, str(step).replace(".", "p"))
        out["y_" + key] = y.numpy().astype(np.float32)
        out["u8_" + key] = ref_utils.tensor2np(y.detach())
    np.savez_compressed(os.path.join(OUT, "chop_steps.npz"), nb=1, seed=21, img_seed=31, **out)
    rec = {}
    for (h, w, p, s, step) in ((40, 56, 32, 4, 1.0), (40, 56, 32, 4, 0.75), (50, 70, 32, 2, 0.75), (80, 64, 32, 1, 0.625),
                               (44, 36, 20, 2, 0.9), (33, 47, 16, 1, 0.95), (130, 90, 200, 1, 0.8)):
        pp = min(h, w, p)
        ys, xs = O.tile_origins(h, pp, step), O.tile_origins(w, pp, step)
        g = torch.Generator().manual_seed(13)
        tiles = torch.rand(len(ys) * len(xs), 3, s * pp, s * pp, generator=g)
        rec["out_%d_%d_%d_%d_%s" % (h, w, p, s, str(step).replace(".", "p"))] = \
            ref_utils.recompose_tensor(tiles, h, w, step=step, scale=s).numpy()
    np.savez_compressed(os.path.join(OUT, "recompose_steps.npz"), **rec)


def i2i_fixtures(td):
    """G8: pix2pix UNet / CycleGAN ResNet generators (BASELINE configs[4]) through the reference's
    get_network and run.Model with the extras run.py applies to these families (run.py:295-361)."""
    from oracle import i2i_oracle as I
    out = {}
    # -- raw module forwards on small configurations: train-mode BatchNorm (pix2pix runs with meval False), eval-mode
    #    BatchNorm, InstanceNorm UNet, batch of two; inputs in [-1, 1]
    for tag, kw, seed, shape, train in (("unet_d5_bn_train", {"type": "unet_256", "num_downs": 5, "ngf": 8}, 41, (2, 3, 32, 64), True),
                                        ("unet_d6_bn_eval", {"type": "unet_256", "num_downs": 6, "ngf": 8}, 42, (1, 3, 64, 128), False),
                                        ("unet_d5_in", {"type": "unet_128", "num_downs": 5, "ngf": 16, "norm_type": "instance"}, 43, (1, 3, 64, 96), True),
                                        ("resnet_b2_in", {"type": "resnet_9blocks", "n_blocks": 2, "ngf": 16}, 44, (2, 3, 40, 52), False),
                                        ("resnet_b1_bn_eval", {"type": "resnet_6blocks", "n_blocks": 1, "ngf": 8, "norm_type": "batch"}, 45, (1, 3, 36, 28), False),
                                        ("resnet_b1_bn_train", {"type": "resnet_6blocks", "n_blocks": 1, "ngf": 8, "norm_type": "batch"}, 46, (2, 3, 24, 32), True)):
        torch.manual_seed(seed)
        net = get_network(get_network_G_config(dict(kw), 1))
        I.randomize_norms(net.state_dict(), seed + 100)
        net.train(train)
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(seed)) * 2 - 1
        with torch.no_grad():
            y = net(x.clone())
        out["y_" + tag] = y.numpy().astype(np.float32)
        out["wsum_" + tag] = np.array([float(v.double().sum()) for k, v in sorted(net.state_dict().items())
                                       if "num_batches" not in k and "running" not in k])
    # -- full-size networks through run.Model as `run.py -a unet_256` / `-a resnet_9blocks -norm` would drive them
    torch.manual_seed(51)
    net = get_network(get_network_G_config({"type": "unet_256"}, 1))
    I.randomize_norms(net.state_dict(), 151)
    path = os.path.join(td, "1x_p2p.pth")
    save_model(net, path)
    model = ref_run.Model(path, "unet_256", None, device=torch.device("cpu"), meval=False, strict=True, chop=False)
    img = ref_utils.linear_resize(image(52, 200, 256), 256)          # run.py:412-413
    out["resized_unet256"] = img
    t = ref_utils.np2tensor(img, normalize=True)
    y = model(t.clone())
    out["y_unet256"] = y.detach().numpy().astype(np.float32)
    out["u8_unet256"] = ref_utils.tensor2np(y.detach(), denormalize=True)

    torch.manual_seed(53)
    net = get_network(get_network_G_config({"type": "resnet_9blocks"}, 1))
    path = os.path.join(td, "1x_cg.pth")
    save_model(net, path)
    model = ref_run.Model(path, "resnet_9blocks", None, device=torch.device("cpu"), meval=True, strict=False, chop=True)
    img = image(54, 64, 96)
    t = ref_utils.np2tensor(img, normalize=True)
    y = model.chop_forward(t.clone(), patch_size=32, step=0.5)        # 3 x 5 tiles of 32 px
    out["y_resnet9_chop32"] = y.detach().numpy().astype(np.float32)
    out["u8_resnet9_chop32"] = ref_utils.tensor2np(y.detach(), denormalize=True)
    y = model(t.clone())                                              # __call__: one 64-px tile
    out["y_resnet9_call"] = y.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "i2i.npz"), **out)


def main():
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "i2i":
        with tempfile.TemporaryDirectory() as td:
            i2i_fixtures(td)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "steps":
        with tempfile.TemporaryDirectory() as td:
            steps_fixtures(td)
        return
    # ---- G1: weights recipe: the oracle's make_state_dict must reproduce the reference init.
    checks = {}
    for scale, nb in ((4, 2), (1, 1), (2, 1), (8, 1), (3, 1)):
        net = ref_net(scale, nb)
        sd = net.state_dict()
        digest = np.array([float(v.double().sum()) for v in sd.values()])
        checks["wsum_s%d_nb%d" % (scale, nb)] = digest
        checks["keys_s%d_nb%d" % (scale, nb)] = np.array(list(sd.keys()))
    np.savez_compressed(os.path.join(OUT, "weights_recipe.npz"), **checks)

    # ---- G2: full 4x nb=23 model on a 64x64 image through run.Model (config 1), fp32 CPU.
    with tempfile.TemporaryDirectory() as td:
        net = ref_net(4, 23)
        path = os.path.join(td, "4x_rand_rrdb.pth")
        save_model(net, path)
        model = ref_run.Model(path, "infer", None, device=torch.device("cpu"), chop=True)
        img = image(0, 64, 64)
        t = ref_utils.np2tensor(img)
        y = model(t)
        u8 = ref_utils.tensor2np(y.detach())
        np.savez_compressed(os.path.join(OUT, "rrdb4x_nb23_64x64.npz"), img_seed=0, h=64, w=64,
                            y=y.numpy().astype(np.float32), u8=u8,
                            arch=model.arch, scale=model.scale, in_nc=model.in_nc, out_nc=model.out_nc)

        # ---- G3: multi-tile chop on small nets (tile geometry + blending), several scales
        for scale, nb, (h, w), patch in ((4, 2, (40, 56), 32), (1, 2, (80, 64), 32), (2, 1, (50, 70), 32),
                                         (3, 1, (36, 30), 200)):
            net = ref_net(scale, nb, seed=1)
            path = os.path.join(td, "%dx_small.pth" % scale)
            save_model(net, path)
            model = ref_run.Model(path, "infer", None, device=torch.device("cpu"), chop=True)
            img = image(3, h, w)
            t = ref_utils.np2tensor(img)
            y = model.chop_forward(t, patch_size=patch, step=0.5)
            np.savez_compressed(os.path.join(OUT, "chop_s%d_nb%d_%dx%d_p%d.npz" % (scale, nb, h, w, patch)),
                                img_seed=3, h=h, w=w, patch=patch, scale=scale, nb=nb, seed=1,
                                y=y.numpy().astype(np.float32), u8=ref_utils.tensor2np(y.detach()))

        steps_fixtures(td)
        i2i_fixtures(td)

        # ---- G4: chain 1x + 4x with -cf (config 3, shrunk): run.py semantics by hand
        n1 = ref_net(1, 1, seed=5)
        n4 = ref_net(4, 1, seed=6)
        p1, p4 = os.path.join(td, "1x_rand_jpeg.pth"), os.path.join(td, "4x_rand_fatal.pth")
        save_model(n1, p1)
        save_model(n4, p4)
        m1 = ref_run.Model(p1, "infer", None, device=torch.device("cpu"), chop=True)
        m4 = ref_run.Model(p4, "infer", None, device=torch.device("cpu"), chop=True)
        img = image(7, 40, 56)
        t = ref_utils.np2tensor(img)
        y = m4(m1(t.clone()))
        u8 = ref_utils.tensor2np(y.detach())
        cf = ref_utils.color_fix(img, u8)
        np.savez_compressed(os.path.join(OUT, "chain_1x4x_cf_40x56.npz"), img_seed=7, h=40, w=56,
                            y=y.numpy().astype(np.float32), u8=u8, cf=cf)

        # ---- G4b: ESRGAN+ (plus=True) is auto-detected from the conv1x1 keys
        net = ref_net(4, 2, seed=9, plus=True)
        path = os.path.join(td, "4x_plus.pth")
        save_model(net, path)
        model = ref_run.Model(path, "infer", None, device=torch.device("cpu"), chop=True)
        img = image(12, 40, 48)
        y = model.chop_forward(ref_utils.np2tensor(img), patch_size=32, step=0.5)
        np.savez_compressed(os.path.join(OUT, "plus_s4_nb2_40x48_p32.npz"), img_seed=12, h=40, w=48, patch=32, seed=9,
                            keys=np.array(list(net.state_dict().keys())),
                            wsum=np.array([float(v.double().sum()) for v in net.state_dict().values()]),
                            y=y.numpy().astype(np.float32), u8=ref_utils.tensor2np(y.detach()))

        # ---- G4c: SRResNet: auto-detected as 'srgan', pixelshuffle upsampler
        for scale, nb, (h, w) in ((4, 3, (40, 48)), (2, 2, (36, 44))):
            torch.manual_seed(13)
            net = get_network(get_network_G_config({"type": "sr_resnet", "nb": nb}, scale)).eval()
            with torch.no_grad():
                list(net.state_dict().values())[-1].fill_(0.5)
            path = os.path.join(td, "%dx_srres.pth" % scale)
            save_model(net, path)
            model = ref_run.Model(path, "infer", None, device=torch.device("cpu"), chop=True)
            img = image(14, h, w)
            y = model.chop_forward(ref_utils.np2tensor(img), patch_size=32, step=0.5)
            np.savez_compressed(os.path.join(OUT, "srresnet_s%d_nb%d_%dx%d_p32.npz" % (scale, nb, h, w)), img_seed=14,
                                h=h, w=w, patch=32, seed=13, scale=scale, nb=nb, arch=model.arch, mscale=model.scale,
                                keys=np.array(list(net.state_dict().keys())),
                                wsum=np.array([float(v.double().sum()) for v in net.state_dict().values()]),
                                y=y.numpy().astype(np.float32), u8=ref_utils.tensor2np(y.detach()))

        ppon_fixtures(td)
        pan_fixtures(td)

    # ---- G5: tile geometry of extract_patches_2d for a list of sizes
    geo = {}
    for (h, w, p) in ((1080, 1920, 200), (720, 1280, 200), (512, 512, 200), (64, 64, 200), (256, 320, 200),
                      (300, 200, 200), (201, 401, 200), (72, 88, 48), (999, 1001, 200)):
        pp = min(h, w, p)
        x = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, h, w)
        patches = ref_utils.extract_patches_2d(x, (pp, pp), [0.5, 0.5], batch_first=True).squeeze(0)
        first = patches[:, 0, 0, 0].long()
        geo["%d_%d_%d" % (h, w, p)] = np.stack([(first // w).numpy(), (first % w).numpy()], 1)
    np.savez_compressed(os.path.join(OUT, "tile_geometry.npz"), **geo)

    # ---- G6: recompose_tensor on random tiles (blend weights), fp32
    rec = {}
    for (h, w, p, s) in ((40, 56, 32, 4), (80, 64, 32, 1), (50, 70, 32, 2), (130, 90, 200, 1)):
        pp = min(h, w, p)
        ys, xs = O.tile_origins(h, pp), O.tile_origins(w, pp)
        g = torch.Generator().manual_seed(11)
        tiles = torch.rand(len(ys) * len(xs), 3, s * pp, s * pp, generator=g)
        rec["out_%d_%d_%d_%d" % (h, w, p, s)] = ref_utils.recompose_tensor(tiles, h, w, step=0.5, scale=s).numpy()
    np.savez_compressed(os.path.join(OUT, "recompose.npz"), **rec)

    # ---- G7: colour fix + colour conversions + np2tensor/tensor2np
    cfd = {}
    for name, (h, w, s) in (("a", (24, 32, 4)), ("b", (37, 29, 2)), ("c", (30, 30, 1)), ("d", (45, 60, 3))):
        lr = image(21, h, w)
        sr = image(22, h * s, w * s)
        # make SR a plausible upscaled version: blur of nearest + noise keeps values spread
        sr = (0.5 * np.repeat(np.repeat(lr, s, 0), s, 1) + 0.5 * sr).astype(np.uint8)
        cfd["lr_" + name] = lr
        cfd["sr_" + name] = sr
        cfd["out_" + name] = ref_utils.color_fix(lr, sr)
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    cfd["srgb2linear_ramp"] = ref_utils.srgb2linear(ramp).astype(np.float32)
    lin = np.linspace(-0.1, 1.1, 4097, dtype=np.float32).reshape(-1, 1, 1)
    cfd["linear2srgb_in"] = lin
    cfd["linear2srgb_out"] = ref_utils.linear2srgb(lin)
    img = image(31, 20, 28)
    t = ref_utils.np2tensor(img)
    cfd["np2tensor_img"] = img
    cfd["np2tensor_out"] = t.numpy()
    y = torch.rand(1, 3, 20, 28, generator=torch.Generator().manual_seed(5)) * 1.4 - 0.2
    cfd["tensor2np_in"] = y.numpy()
    cfd["tensor2np_out"] = ref_utils.tensor2np(y)
    np.savez_compressed(os.path.join(OUT, "color.npz"), **cfd)

    # ---- G8: load-time logic: infer_params / mod2normal / swa2normal / name parsing
    meta = {}
    for scale, nb in ((1, 2), (2, 3), (4, 23), (8, 1)):
        sd = ref_net(scale, nb).state_dict()
        m = ref_run.Model.__new__(ref_run.Model)
        m.arch = "esrgan"
        m.scale = None
        m.in_nc = 3
        m.out_nc = 3
        cfg = m.infer_params(sd)
        meta["infer_s%d_nb%d" % (scale, nb)] = np.array([m.scale, cfg["nb"], cfg["nf"], cfg["in_nc"], cfg["out_nc"],
                                                         int(cfg["plus"]), cfg["upscale"]])
    sd = ref_net(4, 23).state_dict()
    mod = ref_utils.normal2mod(dict(sd))
    meta["mod_keys"] = np.array(list(mod.keys()))
    back = ref_utils.mod2normal(mod)
    meta["mod2normal_keys"] = np.array(list(back.keys()))
    meta["scale_names"] = np.array(["4x_foo.pth", "1x_bar.pth", "2X_baz.pth", "foo.pth", "16x.pth", "x4.pth"])
    meta["scale_vals"] = np.array([-1 if ref_run.get_scale_name(n) is None else ref_run.get_scale_name(n)
                                   for n in meta["scale_names"]])
    np.savez_compressed(os.path.join(OUT, "load_logic.npz"), **meta)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] in ("ppon", "pan"):
        torch.set_num_threads(8)
        with tempfile.TemporaryDirectory() as _td:
            (ppon_fixtures if sys.argv[2] == "ppon" else pan_fixtures)(_td)
    else:
        main()
