"""Network config defaults -- mirror of the reference's utils/defaults.py:3-89 for the ESRGAN,
SRResNet, PPON and PAN families (the others are outside the hot-path scope and raise NotImplementedError)."""

_RRDB_ALIASES = ("rrdb_net", "esrgan", "esrgan-lite")
_SRRESNET_ALIASES = ("sr_resnet", "srresnet", "srgan")
_OTHER_KINDS = ("evsrgan", "mrrdb_net", "mesrgan",
                "unet_net", "unet", "resnet_net", "resnet", "wbcunet", "wbcunet_net")


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
    if kind in _OTHER_KINDS or kind.startswith(("unet_", "p2p_", "resnet_", "cg_")):
        raise NotImplementedError(
            "generator [%s] exists in the reference but is outside the B200 RRDB hot-path scope" % kind)
    raise NotImplementedError("Generator model [%s] not recognized" % kind)
