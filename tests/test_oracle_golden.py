"""Pins oracle/rrdb_oracle.py against outputs of the unmodified reference (tests/golden/*.npz,
produced by tools/make_golden.py in the authoring container). CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden, synth_image
from oracle import rrdb_oracle as O


@pytest.mark.parametrize("scale,nb", [(4, 2), (1, 1), (2, 1), (8, 1), (3, 1)])
def test_weight_recipe_matches_reference_init(scale, nb):
    g = golden("weights_recipe.npz")
    sd = O.make_state_dict(scale=scale, nb=nb, seed=0)
    assert list(sd.keys()) == list(g["keys_s%d_nb%d" % (scale, nb)])
    sums = np.array([float(v.double().sum()) for v in sd.values()])
    np.testing.assert_array_equal(sums, g["wsum_s%d_nb%d" % (scale, nb)])


def test_full_model_64x64_config1():
    g = golden("rrdb4x_nb23_64x64.npz")
    sd = O.make_state_dict(scale=4, nb=23, seed=0)
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    y = O.chop_forward(sd, O.np2tensor(img))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    u8 = O.tensor2np(y)
    assert np.abs(u8.astype(int) - g["u8"].astype(int)).max() <= 1
    assert (u8 != g["u8"]).mean() < 1e-3
    p = O.infer_params(sd)
    assert (p["scale"], p["in_nc"], p["out_nc"]) == (int(g["scale"]), int(g["in_nc"]), int(g["out_nc"]))
    # the comparison must not be vacuous
    assert 0.2 < ((g["u8"] > 0) & (g["u8"] < 255)).mean()


@pytest.mark.parametrize("name", ["chop_s4_nb2_40x56_p32.npz", "chop_s1_nb2_80x64_p32.npz",
                                  "chop_s2_nb1_50x70_p32.npz", "chop_s3_nb1_36x30_p200.npz"])
def test_chop_forward_multi_tile(name):
    g = golden(name)
    sd = O.make_state_dict(scale=int(g["scale"]), nb=int(g["nb"]), seed=int(g["seed"]))
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    y = O.chop_forward(sd, O.np2tensor(img), patch_size=int(g["patch"]))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    assert np.abs(O.tensor2np(y).astype(int) - g["u8"].astype(int)).max() <= 1


def test_esrgan_plus_fixture():
    g = golden("plus_s4_nb2_40x48_p32.npz")
    sd = O.make_state_dict(scale=4, nb=2, seed=int(g["seed"]), plus=True)
    assert list(sd.keys()) == list(g["keys"])
    np.testing.assert_array_equal(np.array([float(v.double().sum()) for v in sd.values()]), g["wsum"])
    assert O.infer_params(sd)["plus"] is True
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    y = O.chop_forward(sd, O.np2tensor(img), patch_size=int(g["patch"]))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", ["srresnet_s4_nb3_40x48_p32.npz", "srresnet_s2_nb2_36x44_p32.npz"])
def test_srresnet_fixture(name):
    g = golden(name)
    scale = int(g["scale"])
    sd = O.make_srresnet_state_dict(scale=scale, nb=int(g["nb"]), seed=int(g["seed"]))
    assert list(sd.keys()) == list(g["keys"])
    np.testing.assert_array_equal(np.array([float(v.double().sum()) for v in sd.values()]), g["wsum"])
    assert str(g["arch"]) == "srgan" and int(g["mscale"]) == scale
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    y = O.chop_forward(sd, O.np2tensor(img), patch_size=int(g["patch"]),
                       forward=lambda t: O.srresnet_forward(sd, t, scale))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", ["ppon_s4_nb1_40x48_p32.npz", "ppon_s2_nb2_36x44_p32.npz"])
def test_ppon_fixture(name):
    """PPON: oracle weights recipe and forward (third output, run.py:191-192) vs the reference."""
    g = golden(name)
    scale = int(g["scale"])
    sd = O.make_ppon_state_dict(scale=scale, nb=int(g["nb"]), seed=int(g["seed"]))
    assert list(sd.keys()) == list(g["keys"])
    np.testing.assert_array_equal(np.array([float(v.double().sum()) for v in sd.values()]), g["wsum"])
    assert str(g["arch"]) == "ppon"
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    x = O.np2tensor(img)
    y = O.chop_forward(sd, x, patch_size=int(g["patch"]), scale=scale, forward=lambda t: O.ppon_forward(sd, t, scale)[2])
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    if "out_p" in g.files:
        for got, key in zip(O.ppon_forward(sd, x, scale), ("out_c", "out_s", "out_p")):
            np.testing.assert_allclose(got.numpy(), g[key], rtol=0, atol=2e-5)


PAN_FIXTURES = ["pan_s4_nb2_40x48_p32.npz", "pan_s2_nb1_36x44_p32.npz", "pan_s3_nb1_24x28_p32.npz",
                "pan_s1_nb1_33x40_p32.npz"]


@pytest.mark.parametrize("name", PAN_FIXTURES)
def test_pan_fixture(name):
    """PAN: oracle weights recipe, whole-image forward and chop_forward vs the reference."""
    g = golden(name)
    scale = int(g["scale"])
    sd = O.make_pan_state_dict(scale=scale, nb=int(g["nb"]), seed=int(g["seed"]), gamma=float(g["gamma"]))
    assert sorted(sd.keys()) == sorted(g["keys"])
    np.testing.assert_array_equal(np.array([float(sd[k].double().sum()) for k in g["keys"]]), g["wsum"])
    assert str(g["arch"]) == "pan"
    x = O.np2tensor(synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"])))
    np.testing.assert_allclose(O.pan_forward(sd, x, scale).numpy(), g["whole"], rtol=0, atol=2e-5)
    y = O.chop_forward(sd, x, patch_size=int(g["patch"]), scale=scale, forward=lambda t: O.pan_forward(sd, t, scale))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    # the attention branch is live in these fixtures (FSA.gamma initialises to 0 in the reference)
    sd0 = dict(sd, **{"FSA.gamma": sd["FSA.gamma"] * 0})
    assert (O.pan_forward(sd0, x, scale) - O.pan_forward(sd, x, scale)).abs().max() > 1e-3


def test_tile_geometry():
    g = golden("tile_geometry.npz")
    for key in g.files:
        h, w, p = (int(v) for v in key.split("_"))
        pp = min(h, w, p)
        want = g[key]
        got = np.array([(y, x) for y in O.tile_origins(h, pp) for x in O.tile_origins(w, pp)])
        np.testing.assert_array_equal(got, want)
    assert len(g["1080_1920_200"]) == 190 and len(g["720_1280_200"]) == 84


def test_recompose_matches_reference():
    g = golden("recompose.npz")
    for key in g.files:
        h, w, p, s = (int(v) for v in key.split("_")[1:])
        pp = min(h, w, p)
        n = len(O.tile_origins(h, pp)) * len(O.tile_origins(w, pp))
        tiles = torch.rand(n, 3, s * pp, s * pp, generator=torch.Generator().manual_seed(11))
        out = O.recompose(tiles, h, w, step=0.5, scale=s)
        np.testing.assert_allclose(out.numpy(), g[key], rtol=0, atol=1e-6)


def _step_key(key):
    """'..._32_4_0p75' -> (ints..., 0.75)"""
    parts = key.split("_")
    return parts[:-1], float(parts[-1].replace("p", "."))


def test_recompose_with_other_steps_matches_reference():
    g = golden("recompose_steps.npz")
    for key in g.files:
        head, step = _step_key(key)
        h, w, p, s = (int(v) for v in head[1:])
        pp = min(h, w, p)
        n = len(O.tile_origins(h, pp, step)) * len(O.tile_origins(w, pp, step))
        tiles = torch.rand(n, 3, s * pp, s * pp, generator=torch.Generator().manual_seed(13))
        out = O.recompose(tiles, h, w, step=step, scale=s)
        np.testing.assert_allclose(out.numpy(), g[key], rtol=0, atol=1e-6)


def test_chop_forward_with_other_steps_matches_reference():
    g = golden("chop_steps.npz")
    for key in [k for k in g.files if k.startswith("y_")]:
        head, step = _step_key(key[2:].replace("step", ""))
        scale = int(head[0][1:])
        h, w = (int(v) for v in head[1].split("x"))
        patch = int(head[2][1:])
        sd = O.make_state_dict(scale=scale, nb=int(g["nb"]), seed=int(g["seed"]))
        img = synth_image(int(g["img_seed"]), h, w)
        y = O.chop_forward(sd, O.np2tensor(img), patch_size=patch, step=step)
        np.testing.assert_allclose(y.numpy(), g[key], rtol=0, atol=2e-5)


def test_chain_and_color_fix():
    g = golden("chain_1x4x_cf_40x56.npz")
    img = synth_image(int(g["img_seed"]), int(g["h"]), int(g["w"]))
    sd1 = O.make_state_dict(scale=1, nb=1, seed=5)
    sd4 = O.make_state_dict(scale=4, nb=1, seed=6)
    y = O.chop_forward(sd4, O.chop_forward(sd1, O.np2tensor(img)))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=2e-5)
    u8 = O.tensor2np(y)
    assert np.abs(u8.astype(int) - g["u8"].astype(int)).max() <= 1
    cf = O.color_fix(img, g["u8"])
    d = np.abs(cf.astype(int) - g["cf"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01


def test_color_functions():
    g = golden("color.npz")
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    np.testing.assert_allclose(O.srgb2linear(ramp), g["srgb2linear_ramp"], rtol=1e-6, atol=0)
    out = O.linear2srgb(g["linear2srgb_in"])
    assert np.abs(out.astype(int) - g["linear2srgb_out"].astype(int)).max() == 0
    for name in "abcd":
        got = O.color_fix(g["lr_" + name], g["sr_" + name])
        d = np.abs(got.astype(int) - g["out_" + name].astype(int))
        assert d.max() <= 1, name
        assert (d > 0).mean() < 0.01, name
    np.testing.assert_array_equal(O.np2tensor(g["np2tensor_img"]).numpy(), g["np2tensor_out"])
    np.testing.assert_array_equal(O.tensor2np(torch.from_numpy(g["tensor2np_in"])), g["tensor2np_out"])


def test_load_logic():
    g = golden("load_logic.npz")
    for scale, nb in ((1, 2), (2, 3), (4, 23), (8, 1)):
        sd = O.make_state_dict(scale=scale, nb=nb, seed=0)
        p = O.infer_params(sd)
        want = g["infer_s%d_nb%d" % (scale, nb)]
        assert [p["scale"], p["nb"], p["nf"], p["in_nc"], p["out_nc"], int(p["plus"]), p["scale"]] == list(want)
    sd = O.make_state_dict(scale=4, nb=23, seed=0)
    mod = {k: None for k in g["mod_keys"]}
    # build a 'new-arch' dict with the reference's key names and map it back
    rev = dict(zip(g["mod_keys"], sd.keys()))
    mod = {k: sd[rev[k]] for k in g["mod_keys"]}
    back = O.mod2normal(mod)
    assert list(back.keys()) == list(g["mod2normal_keys"])
    for k in back:
        assert back[k] is not None


def test_flop_accounting():
    assert O.flop_per_lr_pixel(4, 23, 64) == 35853696
    assert O.flop_per_lr_pixel(1, 23, 64) == 33221376
