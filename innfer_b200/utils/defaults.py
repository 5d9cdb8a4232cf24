"""Network config defaults -- mirror of the reference's utils/defaults.py:3-89 for the ESRGAN,
SRResNet, PPON, PAN, pix2pix UNet and CycleGAN ResNet families (the others raise NotImplementedError)."""

_RRDB_ALIASES = ("rrdb_net", "esrgan", "esrgan-lite")
_SRRESNET_ALIASES = ("sr_resnet", "srresnet", "srgan")
_OTHER_KINDS = ("evsrgan", "mrrdb_net", "mesrgan", "wbcunet", "wbcunet_net")


def get_network_G_config(network_G, scale):
    """Expand an alias (str) or partial dict into the full constructor kwargs of the generator."""
    scale = int(scale)
    if isinstance(network_G, str):
        kind, opts = network_G.lower(), {}
    elif isinstance(network_G, dict):
        opts = network_G
        key = "which_model_G" if "which_model_G" in opts else "type"
        kind = opts.pop(key).lower()
    else:
        raise TypeError("network_G must be a str or a dict")

    if kind in _RRDB_ALIASES:
        lite = kind == "esrgan-lite"
        full = {
            "type": "rrdb_net",
            "norm_type": opts.pop("norm_type", None),
            "mode": opts.pop("mode", "CNA"),
            "nf": opts.pop("nf", 32 if lite else 64),
            "nb": opts.pop("nb", 12 if lite else 23),
            "nr": opts.pop("nr", 3),
            "in_nc": opts.pop("in_nc", 3),
            "out_nc": opts.pop("out_nc", 3),
            "gc": opts.pop("gc", 32),
            "convtype": opts.pop("convtype", "Conv2D"),
            "act_type": opts.pop("net_act", None) or opts.pop("act_type", "leakyrelu"),
            "gaussian_noise": opts.pop("gaussian", True),
            "plus": opts.pop("plus", False),
            "finalact": opts.pop("finalact", None),
            "upscale": opts.pop("scale", scale),
            "upsample_mode": opts.pop("upsample_mode", "upconv"),
        }
        return full
    if kind in _SRRESNET_ALIASES:
        return {
            "type": "sr_resnet",
            "in_nc": opts.pop("in_nc", 3),
            "out_nc": opts.pop("out_nc", 3),
            "nf": opts.pop("nf", 64),
            "nb": opts.pop("nb", 16),
            "upscale": opts.pop("scale", scale),
            "norm_type": opts.pop("norm_type", None),
            "act_type": opts.pop("net_act", None) or opts.pop("act_type", "relu"),
            "mode": opts.pop("mode", "CNA"),
            "upsample_mode": opts.pop("upsample_mode", "pixelshuffle"),
            "convtype": opts.pop("convtype", "Conv2D"),
            "finalact": opts.pop("finalact", None),
            "res_scale": opts.pop("res_scale", 1),
        }
    if "ppon" in kind:
        return {
            "type": "ppon",
            "in_nc": opts.pop("in_nc", 3),
            "out_nc": opts.pop("out_nc", 3),
            "nf": opts.pop("nf", 64),
            "nb": opts.pop("nb", 24),
            "upscale": opts.pop("scale", scale),
            "act_type": opts.pop("net_act", None) or opts.pop("act_type", "leakyrelu"),
            "alpha": opts.pop("alpha", 1),
        }
    if kind in ("pan_net", "pan"):
        return {
            "type": "pan_net",
            "in_nc": opts.pop("in_nc", 3),
            "out_nc": opts.pop("out_nc", 3),
            "nf": opts.pop("nf", 40),
            "unf": opts.pop("unf", 24),
            "nb": opts.pop("nb", 16),
            "scale": opts.pop("scale", scale),
            "self_attention": opts.pop("self_attention", True),
            "double_scpa": opts.pop("double_scpa", False),
            "ups_inter_mode": opts.pop("ups_inter_mode", "nearest"),
        }
    if kind not in _OTHER_KINDS and "wbcunet" not in kind and ("unet" in kind or "p2p" in kind):
        # pix2pix UNet (defaults.py:90-112)
        downs = 7 if kind in ("unet_128", "p2p_128") else 8
        return {
            "type": "unet_net",
            "input_nc": opts.pop("in_nc", 3),
            "output_nc": opts.pop("out_nc", 3),
            "num_downs": opts.pop("num_downs", downs),
            "ngf": opts.pop("ngf", 64),
            "norm_type": opts.pop("norm_type", "batch"),
            "use_dropout": opts.pop("use_dropout", False),
            "upsample_mode": opts.pop("upsample_mode", "deconv"),
        }
    if kind not in _OTHER_KINDS and (("resnet" in kind and kind != "sr_resnet") or "cg" in kind):
        # CycleGAN ResNet (defaults.py:113-131)
        blocks = 6 if kind in ("resnet_6blocks", "resnet_6", "cg_6") else 9
        return {
            "type": "resnet_net",
            "input_nc": opts.pop("in_nc", 3),
            "output_nc": opts.pop("out_nc", 3),
            "n_blocks": opts.pop("n_blocks", blocks),
            "ngf": opts.pop("ngf", 64),
            "norm_type": opts.pop("norm_type", "instance"),
            "use_dropout": opts.pop("use_dropout", False),
            "upsample_mode": opts.pop("upsample_mode", "deconv"),
            "padding_type": opts.pop("padding_type", "reflect"),
        }
    if kind in _OTHER_KINDS or "wbcunet" in kind:
        raise NotImplementedError(
            "generator [%s] exists in the reference but is outside the B200 RRDB hot-path scope" % kind)
    raise NotImplementedError("Generator model [%s] not recognized" % kind)
