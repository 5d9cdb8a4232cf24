"""pix2pix UNet / CycleGAN ResNet generators (BASELINE configs[4]).

CPU tests pin oracle/i2i_oracle.py and the nn.Module mirrors to fixtures written by the unmodified reference
(tests/golden/i2i.npz, tools/make_golden.py i2i); the GPU tests compare the sm_100a engine (csrc/i2i.cu through the
C-ABI) with the same fixtures and the oracle: fp16 mode <= 1/255 on the uint8 image and >= 50 dB, fp32 mode <= 1e-4
relative on the float tensor.
"""
import numpy as np
import pytest
import torch

from conftest import golden, psnr_u8, synth_image
from innfer_b200.architectures import get_network
from innfer_b200.utils import utils as U
from innfer_b200.utils.defaults import get_network_G_config
from oracle import i2i_oracle as I

SMALL = (
    # tag, family, net kwargs, seed, input shape, train-mode norm statistics
    ("unet_d5_bn_train", "unet", dict(num_downs=5, ngf=8, norm="batch"), 41, (2, 3, 32, 64), True),
    ("unet_d6_bn_eval", "unet", dict(num_downs=6, ngf=8, norm="batch"), 42, (1, 3, 64, 128), False),
    ("unet_d5_in", "unet", dict(num_downs=5, ngf=16, norm="instance"), 43, (1, 3, 64, 96), True),
    ("resnet_b2_in", "resnet", dict(n_blocks=2, ngf=16, norm="instance"), 44, (2, 3, 40, 52), False),
    ("resnet_b1_bn_eval", "resnet", dict(n_blocks=1, ngf=8, norm="batch"), 45, (1, 3, 36, 28), False),
    ("resnet_b1_bn_train", "resnet", dict(n_blocks=1, ngf=8, norm="batch"), 46, (2, 3, 24, 32), True),
)


def _state_dict(family, kw, seed):
    sd = I.make_unet_state_dict(seed=seed, **kw) if family == "unet" else I.make_resnet_state_dict(seed=seed, **kw)
    return I.randomize_norms(sd, seed + 100)


def _oracle(family, sd, x, kw, train):
    if family == "unet":
        return I.unet_forward(sd, x, kw["num_downs"], kw["norm"], train)
    return I.resnet_forward(sd, x, kw["n_blocks"], kw["norm"], train)


def _input(seed, shape):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed)) * 2 - 1


def _mirror(family, kw, sd, train):
    cfg = {"type": "unet_256" if family == "unet" else "resnet_9blocks", "ngf": kw["ngf"], "norm_type": kw["norm"]}
    if family == "unet":
        cfg["num_downs"] = kw["num_downs"]
    else:
        cfg["n_blocks"] = kw["n_blocks"]
    net = get_network(get_network_G_config(cfg, 1))
    net.load_state_dict(sd, strict=True)
    return net.train(train)


@pytest.mark.parametrize("tag,family,kw,seed,shape,train", SMALL)
def test_oracle_and_mirror_match_reference_fixture(tag, family, kw, seed, shape, train):
    g = golden("i2i.npz")
    sd = _state_dict(family, kw, seed)
    digest = np.array([float(v.double().sum()) for k, v in sorted(sd.items()) if "num_batches" not in k and "running" not in k])
    np.testing.assert_allclose(digest, g["wsum_" + tag], rtol=0, atol=0)     # same seeded weights as the reference
    x = _input(seed, shape)
    y = _oracle(family, sd, x, kw, train)
    np.testing.assert_allclose(y.numpy(), g["y_" + tag], rtol=0, atol=2e-6)
    net = _mirror(family, kw, sd, train)
    with torch.no_grad():
        ym = net(x.clone())
    np.testing.assert_allclose(ym.numpy(), g["y_" + tag], rtol=0, atol=2e-6)


def test_full_size_networks_through_run_model_cpu(tmp_path):
    """`run.py -a unet_256` / `-a resnet_9blocks` semantics on CPU against the reference: explicit arch, scale 1,
    normalised images, pix2pix in training mode without chop, cyclegan with chop."""
    from innfer_b200 import run as R
    g = golden("i2i.npz")
    sd = I.randomize_norms(I.make_unet_state_dict(seed=51), 151)
    torch.save(sd, tmp_path / "1x_p2p.pth")
    m = R.Model(str(tmp_path / "1x_p2p.pth"), "unet_256", None, device=torch.device("cpu"), meval=False, strict=True, chop=False)
    img = U.linear_resize(synth_image(52, 200, 256), 256)
    np.testing.assert_array_equal(img, g["resized_unet256"])
    y = m(U.np2tensor(img, normalize=True))
    np.testing.assert_allclose(y.detach().numpy(), g["y_unet256"], rtol=0, atol=1e-5)
    np.testing.assert_array_equal(U.tensor2np(y.detach(), denormalize=True), g["u8_unet256"])
    sd = I.make_resnet_state_dict(seed=53)
    torch.save(sd, tmp_path / "1x_cg.pth")
    m = R.Model(str(tmp_path / "1x_cg.pth"), "resnet_9blocks", None, device=torch.device("cpu"), meval=True, strict=False, chop=True)
    t = U.np2tensor(synth_image(54, 64, 96), normalize=True)
    y = m.chop_forward(t, patch_size=32, step=0.5)
    np.testing.assert_allclose(y.detach().numpy(), g["y_resnet9_chop32"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(m(t).detach().numpy(), g["y_resnet9_call"], rtol=0, atol=1e-5)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _u8(y):
    return U.tensor2np(y.detach().float().cpu(), denormalize=True)


@pytest.mark.gpu
@pytest.mark.parametrize("fp16", [True, False])
@pytest.mark.parametrize("tag,family,kw,seed,shape,train", SMALL)
def test_engine_matches_reference_fixture(dev, tag, family, kw, seed, shape, train, fp16):
    g = golden("i2i.npz")
    sd = _state_dict(family, kw, seed)
    net = _mirror(family, kw, sd, train).to(dev)
    if fp16:
        net.half()
    x = _input(seed, shape).to(dev, torch.float16 if fp16 else torch.float32)
    with torch.no_grad():
        y = net(x)
    assert y.dtype == x.dtype and tuple(y.shape) == tuple(g["y_" + tag].shape)
    ref = torch.from_numpy(g["y_" + tag])
    if fp16:
        for b in range(shape[0]):
            a, r = _u8(y[b:b + 1]), _u8(ref[b:b + 1])
            assert np.abs(a.astype(int) - r.astype(int)).max() <= 1, tag
            assert psnr_u8(a, r) >= 50.0, tag
    else:
        assert ((y.float().cpu() - ref).abs().max() / ref.abs().max()).item() <= 1e-4, tag


@pytest.mark.gpu
@pytest.mark.parametrize("fp16", [True, False])
def test_full_size_networks_through_run_model_gpu(dev, tmp_path, fp16):
    """BASELINE configs[4] at 256x256: unet_256 (pix2pix extras) and resnet_9blocks (cyclegan extras, chop) through
    run.Model on the GPU against the reference fixture."""
    from innfer_b200 import run as R
    g = golden("i2i.npz")
    dt = torch.float16 if fp16 else torch.float32
    torch.save(I.randomize_norms(I.make_unet_state_dict(seed=51), 151), tmp_path / "1x_p2p.pth")
    m = R.Model(str(tmp_path / "1x_p2p.pth"), "unet_256", None, device=dev, meval=False, strict=True, chop=False)
    if fp16:
        m.model.half()
    y = m(U.np2tensor(g["resized_unet256"], normalize=True).to(dev, dt))
    if fp16:
        a = _u8(y)
        assert np.abs(a.astype(int) - g["u8_unet256"].astype(int)).max() <= 1
        assert psnr_u8(a, g["u8_unet256"]) >= 50.0
    else:
        assert (np.abs(y.float().cpu().numpy() - g["y_unet256"]).max() / np.abs(g["y_unet256"]).max()) <= 1e-4
    torch.save(I.make_resnet_state_dict(seed=53), tmp_path / "1x_cg.pth")
    m = R.Model(str(tmp_path / "1x_cg.pth"), "resnet_9blocks", None, device=dev, meval=True, strict=False, chop=True)
    if fp16:
        m.model.half()
    t = U.np2tensor(synth_image(54, 64, 96), normalize=True).to(dev, dt)
    for y, key in ((m.chop_forward(t, patch_size=32, step=0.5), "y_resnet9_chop32"), (m(t), "y_resnet9_call")):
        ref = torch.from_numpy(g[key])
        if fp16:
            a, r = _u8(y), _u8(ref)
            assert np.abs(a.astype(int) - r.astype(int)).max() <= 1, key
            assert psnr_u8(a, r) >= 50.0, key
        else:
            assert ((y.float().cpu() - ref).abs().max() / ref.abs().max()).item() <= 1e-4, key


@pytest.mark.gpu
def test_config5_sizes_1024_against_oracle(dev):
    """BASELINE configs[4] at 1024x1024 (a smaller width so that the CPU oracle finishes in seconds): whole-image
    UNet forward (bottleneck 4x4) and ResNet forward, fp16, against the oracle."""
    for family, kw, seed in (("unet", dict(num_downs=8, ngf=8, norm="batch"), 61), ("resnet", dict(n_blocks=2, ngf=8, norm="instance"), 62)):
        sd = _state_dict(family, kw, seed)
        x = _input(seed, (1, 3, 1024, 1024))
        ref = _oracle(family, sd, x, kw, True)
        net = _mirror(family, kw, sd, True).to(dev).half()
        with torch.no_grad():
            y = net(x.to(dev, torch.float16))
        a, r = _u8(y), _u8(ref)
        assert np.abs(a.astype(int) - r.astype(int)).max() <= 1, family
        assert psnr_u8(a, r) >= 50.0, family
