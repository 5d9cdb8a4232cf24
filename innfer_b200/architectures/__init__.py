"""get_network -- mirror of the reference's architectures/__init__.py:5-40 (ESRGAN, SRResNet, PPON, PAN, pix2pix UNet
and CycleGAN ResNet generators; MRRDBNet checkpoints are converted to RRDBNet by run.Model, WBC is out of scope)."""

_OUT_OF_SCOPE = ("mrrdb_net", "wbcunet_net")


def get_network(opt_net):
    """Instantiate the network described by ``opt_net`` (``type`` + constructor kwargs)."""
    kind = opt_net.pop("type").lower()
    if kind == "rrdb_net":
        from . import RRDBNet_arch
        return RRDBNet_arch.RRDBNet(**opt_net)
    if kind == "sr_resnet":
        from . import SRResNet_arch
        return SRResNet_arch.SRResNet(**opt_net)
    if kind == "ppon":
        from . import PPON_arch
        return PPON_arch.PPON(**opt_net)
    if kind == "pan_net":
        from . import PAN_arch
        return PAN_arch.PAN(**opt_net)
    if kind == "unet_net":
        from . import UNet_arch
        return UNet_arch.UnetGenerator(**opt_net)
    if kind == "resnet_net":
        from . import ResNet_arch
        return ResNet_arch.ResnetGenerator(**opt_net)
    if kind in _OUT_OF_SCOPE:
        raise NotImplementedError(
            "Model [%s] exists in the reference but is outside the B200 RRDB hot-path scope "
            "(SURVEY.md section 8f)" % kind)
    raise NotImplementedError("Model [{:s}] not recognized".format(kind))
