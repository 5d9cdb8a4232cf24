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


def _gen_conv(dev, cin, cout, h, w, k, stride=1, pad=1, transposed=False, out_pad=0, reflect=False, bias=True, norm=0, act=0,
              final=False, n=1, fp32=False, seed=0):
    """One layer through innfer_gen_conv (csrc/i2i.cu kernels) and the same layer in torch fp64 on the same operands."""
    import torch.nn.functional as F
    from innfer_b200 import _native as native
    lib = native.load()
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, cin, h, w, generator=g) * 2 - 1
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    fan = cin * k * k / (stride * stride if transposed else 1)
    wgt = (torch.rand(*wshape, generator=g) * 2 - 1) * (2.0 / np.sqrt(fan))
    b = (torch.rand(cout, generator=g) - 0.5) if bias else None
    gam = torch.rand(cout, generator=g) + 0.5
    bet = torch.rand(cout, generator=g) - 0.5
    dt = torch.float32 if fp32 else torch.float16
    xd = x.to(dev, dt)
    xr = xd.double().cpu()
    wr = wgt.double() if fp32 else wgt.half().double()
    br = b.double() if bias else None
    if transposed:
        ref = F.conv_transpose2d(xr, wr, br, stride=stride, padding=pad, output_padding=out_pad)
    elif reflect:
        ref = F.conv2d(F.pad(xr, (pad,) * 4, mode="reflect"), wr, br, stride=stride)
    else:
        ref = F.conv2d(xr, wr, br, stride=stride, padding=pad)
    if norm == 1:
        ref = F.instance_norm(ref, eps=1e-5)
    elif norm == 2:
        ref = F.batch_norm(ref, None, None, gam.double(), bet.double(), training=True, eps=1e-5)
    ref = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.2), 3: torch.tanh}[act](ref)
    y = torch.empty(ref.shape, device=dev, dtype=dt)
    wc = wgt.contiguous().numpy()
    bc = b.contiguous().numpy() if bias else None
    gc, bec = gam.contiguous().numpy(), bet.contiguous().numpy()
    native.check(lib.innfer_gen_conv(xd.data_ptr(), n, cin, h, w, wc.ctypes.data, bc.ctypes.data if bias else None, cout, k,
                                     stride, pad, int(transposed), out_pad, int(reflect), norm,
                                     gc.ctypes.data if norm == 2 else None, bec.ctypes.data if norm == 2 else None, act,
                                     int(final), y.data_ptr(), native.INNFER_F32 if fp32 else native.INNFER_F16, None))
    torch.cuda.synchronize()
    return y.double().cpu(), ref


GEN_CONVS = [
    # UNet: 4x4 stride-2 convolutions (first layer 3 channels; N tiles of 64 / 128; LeakyReLU from the apply kernel)
    dict(cin=3, cout=64, h=64, w=64, k=4, stride=2, bias=False),
    dict(cin=64, cout=128, h=32, w=48, k=4, stride=2, bias=False, norm=2, act=2, n=2),
    dict(cin=16, cout=24, h=20, w=36, k=4, stride=2, norm=1, act=2),
    # inner UNet levels: almost no pixels, long K -> split-K partial tensors
    dict(cin=512, cout=512, h=4, w=4, k=4, stride=2, bias=False, act=1),
    dict(cin=256, cout=256, h=2, w=2, k=4, stride=2, norm=2, act=1, n=3),
    # UNet: 4x4 stride-2 transposed convolutions (four output phases of 2x2 taps)
    dict(cin=1024, cout=512, h=2, w=2, k=4, stride=2, transposed=True, bias=False, norm=2, act=1, n=2),
    dict(cin=128, cout=64, h=16, w=24, k=4, stride=2, transposed=True, bias=False, norm=2, act=1),
    dict(cin=128, cout=3, h=32, w=32, k=4, stride=2, transposed=True, act=3, final=True),
    # ResNet: reflection-padded 7x7 first / last convolutions, stride-2 3x3, reflection-padded 3x3, 3x3 transposed
    dict(cin=3, cout=64, h=40, w=52, k=7, pad=3, reflect=True, norm=1, act=1),
    dict(cin=64, cout=3, h=40, w=52, k=7, pad=3, reflect=True, act=3, final=True),
    dict(cin=64, cout=128, h=40, w=52, k=3, stride=2, norm=1, act=1),
    dict(cin=64, cout=128, h=13, w=17, k=3, stride=2, bias=False),
    dict(cin=256, cout=256, h=16, w=20, k=3, reflect=True, norm=1, act=1, n=2),
    dict(cin=32, cout=32, h=9, w=11, k=3, reflect=True, bias=False, norm=2),
    dict(cin=256, cout=128, h=10, w=13, k=3, stride=2, transposed=True, out_pad=1, norm=1, act=1),
    dict(cin=64, cout=32, h=12, w=12, k=3, stride=2, transposed=True, out_pad=1, bias=False, norm=2, act=1, n=2),
    # more than one M tile per image and per phase, several images
    dict(cin=64, cout=64, h=50, w=70, k=3, reflect=True, act=1, n=3),
    # halo-tile geometry: odd chunk count (second K-chunk of the last slab empty), ragged right / bottom edges, N tile 128 x 2
    dict(cin=24, cout=40, h=37, w=45, k=3, reflect=True, norm=1, act=1),
    dict(cin=40, cout=256, h=70, w=41, k=4, stride=2, bias=False, norm=2, act=2),
    dict(cin=136, cout=72, h=21, w=19, k=4, stride=2, transposed=True, norm=2, act=1, n=2),
    dict(cin=8, cout=8, h=33, w=130, k=7, pad=3, reflect=True, act=3, final=True, n=2),
    # under one wave of halo tiles with a long K: the slabs of a tile split over two CTAs (fp32 partial tensors)
    dict(cin=128, cout=256, h=64, w=96, k=3, reflect=True, norm=1, act=1),
    dict(cin=256, cout=256, h=64, w=64, k=4, stride=2, bias=False, norm=2, act=2, n=2),
]


# layers of GEN_CONVS whose geometry the halo-tile kernel takes (at least 8 x 8 output pixels per phase)
HALO_CASES = (0, 1, 2, 6, 7, 8, 9, 10, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["im2col", "halo"])
@pytest.mark.parametrize("cfg", GEN_CONVS, ids=lambda c: "-".join("%s%s" % (k, v) for k, v in c.items()))
def test_generator_layer_tcgen05(dev, cfg, kernel, monkeypatch):
    """Every layer type of the two generators on both tcgen05 kernels -- im2col gather with split-K, and the halo-tile
    variant (parity planes for stride 2, phases for transposed convolutions) -- + the normalisation kernels against
    torch fp64 on the same fp16-rounded operands."""
    from innfer_b200 import _native as native
    idx = GEN_CONVS.index(cfg)
    if kernel == "halo" and idx not in HALO_CASES:
        pytest.skip("too few output pixels for 16 x 8 patches: served by the im2col kernel")
    monkeypatch.setenv("INNFER_I2I_HALO", "2" if kernel == "halo" else "0")
    n0 = native.load().innfer_debug_i2i_halo_launches()
    y, ref = _gen_conv(dev, **cfg)
    assert (native.load().innfer_debug_i2i_halo_launches() - n0 == 1) == (kernel == "halo")
    assert torch.isfinite(y).all()
    tol = 2e-3 * max(1.0, ref.abs().max().item())
    assert (y - ref).abs().max().item() <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [c for i, c in enumerate(GEN_CONVS) if i in (0, 2, 5, 7, 8, 9, 13, 15)],
                         ids=lambda c: "-".join("%s%s" % (k, v) for k, v in c.items()))
def test_generator_layer_fp32_kernel(dev, cfg):
    y, ref = _gen_conv(dev, fp32=True, **cfg)
    assert ((y - ref).abs().max() / ref.abs().max()).item() <= 1e-5


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


@pytest.mark.gpu
@pytest.mark.parametrize("arch,shape", [("resnet_9blocks", (72, 100)), ("unet_256", (200, 256))])
def test_cli_gpu_matches_cli_cpu(dev, tmp_path, monkeypatch, arch, shape):
    """`python run.py -m <model> -a <arch>` on the GPU (fp16 engine) against the same command with -cpu (torch modules,
    fp32): pix2pix resizes to a multiple of 256 and runs the whole image in training mode, cyclegan chops."""
    import cv2
    from innfer_b200 import run as R
    (tmp_path / "models").mkdir()
    (tmp_path / "input").mkdir()
    for d in ("out_gpu", "out_cpu"):
        (tmp_path / d).mkdir()
    sd = (I.randomize_norms(I.make_unet_state_dict(seed=71), 171) if arch == "unet_256" else I.make_resnet_state_dict(seed=72))
    torch.save(sd, tmp_path / "models" / "1x_rand_i2i.pth")
    cv2.imwrite(str(tmp_path / "input" / "a.png"), synth_image(73, *shape))
    monkeypatch.chdir(tmp_path)
    R.main(["-m", "i2i", "-a", arch, "-i", "input", "-o", "out_gpu"])
    R.main(["-m", "i2i", "-a", arch, "-i", "input", "-o", "out_cpu", "-cpu"])
    a = cv2.imread(str(tmp_path / "out_gpu" / "a.png"), cv2.IMREAD_UNCHANGED)
    b = cv2.imread(str(tmp_path / "out_cpu" / "a.png"), cv2.IMREAD_UNCHANGED)
    assert a.shape == b.shape and a.shape[0] % 4 == 0
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1
    assert psnr_u8(a, b) >= 50.0


@pytest.mark.gpu
@pytest.mark.parametrize("family,kw,seed,shape", [("unet", dict(num_downs=5, ngf=8, norm="batch"), 81, (1, 3, 64, 96)),
                                                  ("resnet", dict(n_blocks=2, ngf=16, norm="instance"), 82, (2, 3, 40, 52))])
def test_graph_replay_is_bit_identical(dev, family, kw, seed, shape):
    """Small forwards are recorded into a CUDA graph the second time a (buffers, shape) combination is seen: the eager
    first call, the recording call and the replays must give the same bits, for a new input in the same buffers too."""
    from innfer_b200 import _native as native
    lib = native.load()
    sd = _state_dict(family, kw, seed)
    net = _mirror(family, kw, sd, True).to(dev).half()
    x = _input(seed, shape).to(dev, torch.float16)
    n0 = lib.innfer_debug_i2i_graph_replays()
    with torch.no_grad():
        ys = [net(x).clone() for _ in range(4)]
        y_other = net(-x)
        y_back = net(x)
    assert lib.innfer_debug_i2i_graph_replays() - n0 >= 4
    for y in ys[1:] + [y_back]:
        assert torch.equal(y, ys[0])
    assert not torch.equal(y_other, ys[0])
    ref = _oracle(family, sd, _input(seed, shape), kw, True)
    assert (ys[0].float().cpu() - ref).abs().max().item() < 0.02
