"""CPU tests of the host-side mirror (innfer_b200.run / architectures / utils) against fixtures
generated from the unmodified reference, and of the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, synth_image
from innfer_b200 import _native
from innfer_b200 import run as R
from innfer_b200.architectures import get_network
from innfer_b200.utils import utils as U
from innfer_b200.utils.defaults import get_network_G_config
from oracle import rrdb_oracle as O


def _save(sd, path):
    torch.save(sd, path)
    return str(path)


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "innfer_b200.h")).read()
    declared = set(re.findall(r"\b(innfer_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _native.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_native.EXPORTED_SYMBOLS)


def test_tiles_plan_matches_reference_geometry():
    lib = _native.load()
    g = golden("tile_geometry.npz")
    for key in g.files:
        h, w, p = (int(v) for v in key.split("_"))
        tiles = (_native.Tile * 4096)()
        n, ts = ctypes.c_int(), ctypes.c_int()
        _native.check(lib.innfer_tiles_plan(h, w, p, 0.5, tiles, 4096, ctypes.byref(n), ctypes.byref(ts)))
        got = np.array([(tiles[i].y0, tiles[i].x0) for i in range(n.value)])
        np.testing.assert_array_equal(got, g[key])
        assert ts.value == min(h, w, p)


def test_abi_rejects_bad_arguments_without_gpu():
    lib = _native.load()
    n = ctypes.c_int()
    assert lib.innfer_tiles_plan(0, 10, 200, 0.5, None, 0, ctypes.byref(n), None) == -1
    assert b"tile" in lib.innfer_last_error()
    cfg = _native.RRDBCfg(3, 3, 48, 23, 32, 4, 0, 1)  # nf=48 is not implemented: must fail loudly
    h = ctypes.c_void_p()
    assert lib.innfer_rrdb_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -2
    assert b"nf" in lib.innfer_last_error()
    # PAN: nf must be a multiple of 8 (<= 64); the bilinear skip needs in_nc == out_nc
    pan = _native.PANCfg(3, 3, 44, 24, 16, 4, 1, 0, 1)
    assert lib.innfer_pan_create(ctypes.byref(pan), 0, ctypes.byref(h)) == -2 and b"nf" in lib.innfer_last_error()
    pan = _native.PANCfg(3, 1, 40, 24, 16, 4, 1, 0, 1)
    assert lib.innfer_pan_create(ctypes.byref(pan), 0, ctypes.byref(h)) == -2 and b"in_nc" in lib.innfer_last_error()


def test_state_dict_keys_and_cpu_forward_match_reference(tmp_path):
    g = golden("rrdb4x_nb23_64x64.npz")
    sd = O.make_state_dict(scale=4, nb=23, seed=0)
    path = _save(sd, tmp_path / "4x_rand_rrdb.pth")
    m = R.Model(path, "infer", None, device=torch.device("cpu"), chop=True)
    assert (m.arch, m.scale, m.in_nc, m.out_nc) == (str(g["arch"]), int(g["scale"]), int(g["in_nc"]), int(g["out_nc"]))
    assert list(m.model.state_dict().keys()) == list(sd.keys())
    img = synth_image(0, 64, 64)
    y = m(U.np2tensor(img))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    assert np.abs(U.tensor2np(y).astype(int) - g["u8"].astype(int)).max() <= 1


@pytest.mark.parametrize("name", ["chop_s4_nb2_40x56_p32.npz", "chop_s1_nb2_80x64_p32.npz", "chop_s2_nb1_50x70_p32.npz"])
def test_cpu_mode_chop_forward(tmp_path, name):
    g = golden(name)
    scale = int(g["scale"])
    sd = O.make_state_dict(scale=scale, nb=int(g["nb"]), seed=int(g["seed"]))
    m = R.Model(_save(sd, tmp_path / ("%dx_t.pth" % scale)), "infer", None, device=torch.device("cpu"))
    assert m.scale == scale
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    y = m.chop_forward(U.np2tensor(img), patch_size=int(g["patch"]), step=0.5)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)


def test_esrgan_plus_is_detected_and_runs_on_cpu(tmp_path):
    g = golden("plus_s4_nb2_40x48_p32.npz")
    sd = O.make_state_dict(scale=4, nb=2, seed=int(g["seed"]), plus=True)
    m = R.Model(_save(sd, tmp_path / "4x_plus.pth"), "infer", None, device=torch.device("cpu"))
    assert m.model.cfg["plus"] is True and list(m.model.state_dict().keys()) == list(g["keys"])
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    y = m.chop_forward(U.np2tensor(img), patch_size=int(g["patch"]), step=0.5)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)


def test_srresnet_is_detected_and_runs_on_cpu(tmp_path):
    g = golden("srresnet_s4_nb3_40x48_p32.npz")
    sd = O.make_srresnet_state_dict(scale=4, nb=3, seed=int(g["seed"]))
    m = R.Model(_save(sd, tmp_path / "4x_srres.pth"), "infer", None, device=torch.device("cpu"))
    assert (m.arch, m.scale) == ("srgan", 4) and list(m.model.state_dict().keys()) == list(g["keys"])
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    y = m.chop_forward(U.np2tensor(img), patch_size=int(g["patch"]), step=0.5)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)


def test_ppon_module_mirror_on_cpu(tmp_path):
    """The PPON module tree carries the reference's key names and reproduces its three outputs; run.Model probes
    'CFEM.0.weight' and, like the reference (run.py:157-163), builds the default depth (nb = 24) for it."""
    g = golden("ppon_s2_nb2_36x44_p32.npz")
    sd = O.make_ppon_state_dict(scale=2, nb=2, seed=int(g["seed"]))
    net = get_network(get_network_G_config({"type": "ppon", "nb": 2}, 2)).eval()
    assert list(net.state_dict().keys()) == list(g["keys"])
    net.load_state_dict(sd, strict=True)
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    with torch.no_grad():
        outs = net(U.np2tensor(img))
    for got, key in zip(outs, ("out_c", "out_s", "out_p")):
        np.testing.assert_allclose(got.numpy(), g[key], rtol=0, atol=2e-5)
    m = R.Model.__new__(R.Model)
    m.arch, m.scale, m.model, m.chop = "ppon", 2, net, True
    y = m.chop_forward(U.np2tensor(img), patch_size=int(g["patch"]), step=0.5)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    assert get_network_G_config("ppon", 4) == {"type": "ppon", "in_nc": 3, "out_nc": 3, "nf": 64, "nb": 24, "upscale": 4,
                                               "act_type": "leakyrelu", "alpha": 1}
    with pytest.raises(RuntimeError):   # a 2-block checkpoint does not fit the default 24-block network (strict load)
        R.Model(_save(sd, tmp_path / "2x_ppon.pth"), "infer", 2, device=torch.device("cpu"))


@pytest.mark.parametrize("name", ["pan_s4_nb2_40x48_p32.npz", "pan_s2_nb1_36x44_p32.npz"])
def test_pan_module_mirror_on_cpu(name, tmp_path):
    """The PAN module tree carries the reference's key names (including the upsampler indices that depend on how
    block.sequential flattens: five entries per stage at scale 4, six at scale 2) and reproduces its output; run.Model
    probes 'SCPA_trunk.0.conv1_a.weight' and builds the default depth (nb = 16) for it (run.py:50-53,157-163)."""
    g = golden(name)
    scale, nb = int(g["scale"]), int(g["nb"])
    sd = O.make_pan_state_dict(scale=scale, nb=nb, seed=int(g["seed"]), gamma=float(g["gamma"]))
    net = get_network(get_network_G_config({"type": "pan", "nb": nb}, scale)).eval()
    assert list(net.state_dict().keys()) == list(g["keys"])
    net.load_state_dict(sd, strict=True)
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    with torch.no_grad():
        np.testing.assert_allclose(net(U.np2tensor(img)).numpy(), g["whole"], rtol=0, atol=2e-5)
    m = R.Model.__new__(R.Model)
    m.arch, m.scale, m.model, m.chop = "pan", scale, net, True
    y = m.chop_forward(U.np2tensor(img), patch_size=int(g["patch"]), step=0.5)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    assert get_network_G_config("pan", 4) == {
        "type": "pan_net", "in_nc": 3, "out_nc": 3, "nf": 40, "unf": 24, "nb": 16, "scale": 4, "self_attention": True,
        "double_scpa": False, "ups_inter_mode": "nearest"}
    with pytest.raises(RuntimeError):   # a shallow checkpoint does not fit the default 16-block network (strict load)
        R.Model(_save(sd, tmp_path / ("%dx_pan.pth" % scale)), "infer", scale, device=torch.device("cpu"))


def test_infer_params_and_key_mapping(tmp_path):
    g = golden("load_logic.npz")
    for scale, nb in ((1, 2), (2, 3), (4, 23), (8, 1)):
        sd = O.make_state_dict(scale=scale, nb=nb, seed=0)
        m = R.Model.__new__(R.Model)
        m.arch, m.scale, m.in_nc, m.out_nc = "esrgan", None, 3, 3
        cfg = m.infer_params(sd)
        got = [m.scale, cfg["nb"], cfg["nf"], cfg["in_nc"], cfg["out_nc"], int(cfg["plus"]), cfg["upscale"]]
        assert got == list(g["infer_s%d_nb%d" % (scale, nb)])
    sd = O.make_state_dict(scale=4, nb=23, seed=0)
    mod = U.normal2mod(dict(sd))
    assert list(mod.keys()) == list(g["mod_keys"])
    back = U.mod2normal(mod)
    assert list(back.keys()) == list(g["mod2normal_keys"])
    for k in sd:
        assert torch.equal(back[k], sd[k])
    # a 'new-arch' checkpoint is auto-detected, converted and loads strictly
    m = R.Model(_save(mod, tmp_path / "4x_newarch.pth"), "infer", None, device=torch.device("cpu"))
    assert m.arch == "esrgan" and m.scale == 4
    # SWA wrapper
    swa = {"n_averaged": torch.tensor(3)}
    swa.update({"module.module." + k: v for k, v in O.make_state_dict(scale=1, nb=1).items()})
    assert list(U.swa2normal(swa).keys()) == list(O.make_state_dict(scale=1, nb=1).keys())
    m = R.Model(_save(swa, tmp_path / "1x_swa.pth"), "infer", None, device=torch.device("cpu"))
    assert m.scale == 1
    names, vals = g["scale_names"], g["scale_vals"]
    for n, v in zip(names, vals):
        got = R.get_scale_name(str(n))
        assert (-1 if got is None else got) == int(v)


def test_model_path_resolution(tmp_path, monkeypatch):
    (tmp_path / "models").mkdir()
    for name in ("4x_rand_rrdb.pth", "1x_rand_jpeg.pth", "4x_other_rrdb.pth"):
        torch.save({}, tmp_path / "models" / name)
    monkeypatch.chdir(tmp_path)
    chain, scales = R.parse_models("jpeg+rand_rrdb")
    assert [os.path.basename(c) for c in chain] == ["1x_rand_jpeg.pth", "4x_rand_rrdb.pth"]
    assert scales == [1, 4]
    chain, _ = R.parse_models("JPEG>other")
    assert os.path.basename(chain[1]) == "4x_other_rrdb.pth"
    with pytest.raises(ValueError):
        R.parse_models("rrdb")           # ambiguous filter
    with pytest.raises(IndexError):
        R.parse_models("nosuchmodel")    # the reference falls through to m_list[0]
    with pytest.raises(ValueError):
        R.parse_models("jpeg+rand_rrdb", scales_list=[4])
    assert R.check_model_path("4x_rand_rrdb.pth", None) == os.path.join("models", "4x_rand_rrdb.pth")


def test_synth_recipe_matches_reference_init():
    from innfer_b200 import synth
    for scale, nb in ((4, 2), (1, 1), (8, 1)):
        a, b = synth.make_state_dict(scale=scale, nb=nb, seed=3), O.make_state_dict(scale=scale, nb=nb, seed=3)
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)
    assert synth.flop_per_lr_pixel(4, 23, 64) == 35853696 == O.flop_per_lr_pixel(4, 23, 64)


def test_unknown_architectures_fail_loudly():
    with pytest.raises(NotImplementedError):
        get_network({"type": "wbcunet_net"})
    with pytest.raises(NotImplementedError):
        get_network({"type": "mrrdb_net"})
    with pytest.raises(Exception, match="Could not infer"):
        m = R.Model.__new__(R.Model)
        m.__dict__.update(model_path="x", arch="infer", scale=None, in_nc=3, out_nc=3, device="cpu", eval=True,
                          strict=True, chop=True)
        torch.save({"weird.weight": torch.zeros(1)}, "/tmp/_innfer_weird.pth")
        m.model_path = "/tmp/_innfer_weird.pth"
        m.load_model()
    with pytest.raises(NotImplementedError):   # PAN: only the default 'nearest' upsampler is built
        get_network(get_network_G_config({"type": "pan", "ups_inter_mode": "bilinear"}, 4))
    cfg = get_network_G_config("esrgan-lite", 2)
    assert (cfg["nf"], cfg["nb"], cfg["upscale"], cfg["type"]) == (32, 12, 2, "rrdb_net")


def test_np2tensor_tensor2np_colorfix_host():
    g = golden("color.npz")
    np.testing.assert_array_equal(U.np2tensor(g["np2tensor_img"]).numpy(), g["np2tensor_out"])
    np.testing.assert_array_equal(U.tensor2np(torch.from_numpy(g["tensor2np_in"])), g["tensor2np_out"])
    for name in "abcd":
        np.testing.assert_array_equal(U.color_fix(g["lr_" + name], g["sr_" + name]), g["out_" + name])
    rec = golden("recompose.npz")
    for key in rec.files:
        h, w, p, s = (int(v) for v in key.split("_")[1:])
        pp = min(h, w, p)
        n = len(O.tile_origins(h, pp)) * len(O.tile_origins(w, pp))
        tiles = torch.rand(n, 3, s * pp, s * pp, generator=torch.Generator().manual_seed(11))
        np.testing.assert_allclose(U.recompose_tensor(tiles, h, w, step=0.5, scale=s).numpy(), rec[key], atol=1e-6)


def test_recompose_tensor_other_steps_and_batches():
    """CPU recompose_tensor against reference fixtures for step != 0.5, and final_batch_size > 1 (utils.py:412)."""
    rec = golden("recompose_steps.npz")
    for key in rec.files:
        parts = key.split("_")
        h, w, p, s = (int(v) for v in parts[1:5])
        step = float(parts[5].replace("p", "."))
        pp = min(h, w, p)
        n = len(O.tile_origins(h, pp, step)) * len(O.tile_origins(w, pp, step))
        tiles = torch.rand(n, 3, s * pp, s * pp, generator=torch.Generator().manual_seed(13))
        got = U.recompose_tensor(tiles, h, w, step=step, scale=s)
        np.testing.assert_allclose(got.numpy(), rec[key], atol=1e-6)
        two = U.recompose_tensor(torch.cat([tiles, tiles.flip(0)], 0), h, w, step=step, scale=s)
        assert two.shape[0] == 2
        np.testing.assert_allclose(two[0].numpy(), rec[key][0], atol=1e-6)


def test_engine_cache_follows_the_parameters_and_is_not_pickled():
    """The native engine of a mirror module is keyed on the parameters' storage and version counters and is left
    out of deep copies / pickles.  No GPU needed: the fingerprint logic is host code."""
    import copy
    import pickle
    net = get_network(get_network_G_config({"type": "esrgan", "nb": 1}, 4))
    fp0 = net._fingerprint()
    with torch.no_grad():
        next(net.parameters()).mul_(1.0)          # in-place edit: version counter moves
    assert net._fingerprint() != fp0
    fp1 = net._fingerprint()
    net.load_state_dict(net.state_dict())
    assert net._fingerprint() != fp1
    fp2 = net._fingerprint()
    net.half()
    assert net._fingerprint() != fp2

    class FakeEngine:
        closed = False

        def close(self):
            self.closed = True
    fake = FakeEngine()
    net.__dict__["_engines"] = {("cuda:0", torch.float16): (net._fingerprint(), fake)}
    clone = copy.deepcopy(net)
    assert clone.__dict__["_engines"] == {} and not fake.closed
    again = pickle.loads(pickle.dumps(net))
    assert again.__dict__["_engines"] == {}
    assert set(again.state_dict()) == set(net.state_dict())
    net.invalidate_engine()
    assert fake.closed and net.__dict__["_engines"] == {}


def test_native_chain_needs_cuda_and_three_channel_native_models(tmp_path):
    """run.native_chain: the fused device loop is only taken for CUDA + chop + 3-channel native models."""
    torch.save(O.make_state_dict(scale=4, nb=1, seed=6), tmp_path / "4x_a.pth")
    m = R.Model(str(tmp_path / "4x_a.pth"), "infer", None, device=torch.device("cpu"))
    assert R.native_chain([m], torch.device("cpu"), True) is None


def test_cli_cpu_end_to_end(tmp_path, monkeypatch):
    """python run.py -m jpeg+fatal -cf -cpu on a tiny image (config 3, shrunk)."""
    import cv2
    g = golden("chain_1x4x_cf_40x56.npz")
    (tmp_path / "models").mkdir()
    (tmp_path / "input").mkdir()
    (tmp_path / "output").mkdir()
    torch.save(O.make_state_dict(scale=1, nb=1, seed=5), tmp_path / "models" / "1x_rand_jpeg.pth")
    torch.save(O.make_state_dict(scale=4, nb=1, seed=6), tmp_path / "models" / "4x_rand_fatal.pth")
    cv2.imwrite(str(tmp_path / "input" / "a.png"), synth_image(7, 40, 56))
    monkeypatch.chdir(tmp_path)
    R.main(["-m", "jpeg+fatal", "-cf", "-cpu", "-i", "input", "-o", "output"])
    out = cv2.imread(str(tmp_path / "output" / "a.png"), cv2.IMREAD_UNCHANGED)
    d = np.abs(out.astype(int) - g["cf"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    with pytest.raises(NotImplementedError):
        R.main(["-m", "jpeg", "-a", "wbcunet", "-cpu"])


def test_cuda_request_never_falls_back(tmp_path, monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    (tmp_path / "models").mkdir()
    torch.save(O.make_state_dict(scale=1, nb=1), tmp_path / "models" / "1x_a.pth")
    monkeypatch.chdir(tmp_path)
    with pytest.raises(RuntimeError, match="CUDA is not available"):
        R.main(["-m", "1x_a"])
    from innfer_b200.engine import RRDBEngine
    with pytest.raises(RuntimeError):
        RRDBEngine(dict(in_nc=3, out_nc=3, nf=64, nb=1, scale=1), "cuda:0")
