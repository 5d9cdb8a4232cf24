"""Drop-in mirror of the reference's run.py for the RRDB/ESRGAN path: the ``Model`` wrapper
(run.py:23-225), model-path resolution (229-293) and the CLI (321-443) with the same flags.

Differences that matter: on a CUDA device ``Model.__call__`` / ``chop_forward`` hand the whole
tile loop (extract -> forward -> blend) to the sm_100a engine in ONE native call instead of a
Python loop with a cache flush per tile; on ``-cpu`` the torch modules run as in the reference.
A CUDA request never degrades to the CPU path: errors propagate.
"""
import argparse
import os.path as osp

import torch

from .architectures import get_network
from .utils.defaults import get_network_G_config
from .utils.utils import (color_fix, color_fix_device, extract_patches_2d, get_images_paths, get_models_paths,
                          linear_resize, mod2normal, np2tensor, read_img, recompose_tensor, save_img, save_img_comp, swa2normal, tensor2np)

# key that identifies each architecture family, probed in the reference's order (run.py:50-72)
_ARCH_PROBES = (
    ("SCPA_trunk.0.conv1_a.weight", "pan"),
    ("model.1.sub.0.res.0.weight", "srgan"),
    ("conv_first.weight", "mesrgan"),
    ("model.0.weight", "esrgan"),
    ("CFEM.0.weight", "ppon"),
    ("conv_9.weight", "wbcunet"),
)


class Model:
    def __init__(self, model_path, arch=None, scale=None, in_nc=3, out_nc=3, device="cpu", meval=True,
                 strict=True, chop=True):
        self.model_path = model_path
        self.arch = arch
        self.scale = scale
        self.in_nc = in_nc
        self.out_nc = out_nc
        self.device = device
        self.model = None
        self.eval = meval
        self.strict = strict
        self.chop = chop
        self.load_model()

    # ------------------------------------------------------------------ loading
    def load_model(self):
        if self.arch == "ts":
            # opaque traced graph: cannot sit behind the custom kernels, delegated to torch
            self.model = torch.jit.load(osp.join(self.model_path)).eval().to(self.device)
            return
        state_dict = torch.load(self.model_path, map_location="cpu")
        if "n_averaged" in state_dict:
            state_dict = swa2normal(state_dict)

        if self.arch == "infer":
            for key, arch in _ARCH_PROBES:
                if key in state_dict:
                    self.arch = arch
                    break
            else:
                raise Exception("Could not infer model parameters.")
            if self.arch == "mesrgan":
                state_dict = mod2normal(state_dict)
                self.arch = "esrgan"
            net_params = self.infer_params(state_dict)
        else:
            if not self.scale:
                self.scale = 1
            net_params = get_network_G_config({"type": self.arch}, self.scale)

        net = get_network(net_params)
        net.load_state_dict(state_dict, strict=self.strict)
        del state_dict
        for _, v in net.named_parameters():
            v.requires_grad = False
        if self.eval:
            net.eval()
        self.model = net.to(self.device)

    def infer_params(self, state_dict):
        """Read nb / nf / in_nc / out_nc / scale / plus off the state-dict keys (run.py:103-165)."""
        if self.arch in ("esrgan", "srgan"):
            n_2x, top_index, out_nc, nb, plus = 0, 0, None, None, False
            for key in list(state_dict):
                parts = key.split(".")
                if len(parts) == 5 and parts[2] == "sub":
                    nb = int(parts[3])
                elif len(parts) == 3:
                    index = int(parts[1])
                    if index > 6 and parts[0] == "model" and parts[2] == "weight":
                        n_2x += 1
                    if index > top_index:
                        top_index = index
                        out_nc = state_dict[key].shape[0]
                if self.arch == "esrgan" and "conv1x1" in key:
                    plus = True
            self.in_nc = state_dict["model.0.weight"].shape[1]
            self.out_nc = out_nc
            self.scale = 2 ** n_2x
            net_dict = {"type": self.arch, "in_nc": self.in_nc, "out_nc": self.out_nc,
                        "nf": state_dict["model.0.weight"].shape[0], "nb": nb}
            if self.arch == "esrgan":
                net_dict["plus"] = plus
        elif self.arch == "wbcunet":
            self.scale = 1
            net_dict = {"type": self.arch, "mode": "pt", "nf": state_dict["conv.weight"].shape[0]}
        else:  # ppon, pan
            net_dict = {"type": self.arch, "in_nc": self.in_nc, "out_nc": self.out_nc}
        return get_network_G_config(net_dict, self.scale)

    # ------------------------------------------------------------------ inference
    def chop_forward(self, data, patch_size=200, step=1.0):
        """Tile the image into (patch_size, patch_size) crops, run the network per crop and blend."""
        _, _, height, width = data.size()
        patch_size = min(height, width, patch_size)
        if data.is_cuda and hasattr(self.model, "chop_forward_native"):
            return self.model.chop_forward_native(data, patch_size, step)
        patches = extract_patches_2d(img=data, patch_shape=(patch_size, patch_size), step=[step, step],
                                     batch_first=True).squeeze(0)
        outputs = []
        with torch.no_grad():
            for i in range(patches.size(0)):
                pred = self.model(patches[i:i + 1])
                if self.arch == "ppon":
                    pred = pred[2]
                if self.arch == "ts":
                    pred = pred.detach().cpu()
                outputs.append(pred)
        return recompose_tensor(torch.cat(outputs, 0), height, width, step=step, scale=self.scale)

    def __call__(self, data):
        if self.chop:
            return self.chop_forward(data=data, patch_size=200, step=0.5)
        with torch.no_grad():
            out = self.model(data)
        return out[2] if self.arch == "ppon" else out


# ---------------------------------------------------------------------- model path handling
def parse_models(models_paths, scales_list=None):
    """'a+b' (or 'a>b') -> ([resolved paths], [scales parsed from the file names])."""
    chain = models_paths.split("+") if "+" in models_paths else models_paths.split(">")
    available = get_models_paths("./models")
    full_chain = [check_model_path(m, available) for m in chain]
    if not scales_list:
        scales_list = [get_scale_name(m, None) for m in full_chain]
    elif len(scales_list) != len(chain):
        raise ValueError(f"The num. of scales {len(scales_list)} is != from number of models {len(chain)}")
    return full_chain, scales_list


def check_model_path(model_path, all_models=None):
    """Exact path, then ./models/<name>, then case-insensitive substring match over ./models."""
    if osp.isfile(model_path):
        return model_path
    in_models = osp.join("models", model_path)
    if osp.isfile(in_models):
        return in_models
    if not all_models:
        raise ValueError(f"Model {model_path} not found.")
    hits = [m for m in all_models if str(model_path.lower()) in str(m).lower()]
    if len(hits) > 1:
        raise ValueError(f"Filter {model_path} returned multiple models: {hits}.")
    return hits[0]  # IndexError when nothing matches, like the reference


def get_scale_name(model_path, scale=None):
    """Scale from the first two characters of the file name ('4x_...')."""
    found = None
    head = str(osp.basename(model_path)[0:2]).lower()
    if "x" in head:
        try:
            found = int(head.replace("x", ""))
        except ValueError:
            found = None
    if scale:
        if found and scale != found:
            print(f"Warning: possible model scale mismatch on {model_path}")
        return scale
    return found


default_extras = {"meval": True, "strict": True, "normalize": False}
# run.py:299-309 of the reference: pix2pix keeps its norm layers in training mode and, like cyclegan, works on images
# normalised to [-1, 1]; cyclegan loads non-strictly (running statistics saved by PyTorch < 0.4 InstanceNorm layers)
pix2pix_extras = {"meval": False, "strict": True, "normalize": True}
cyclegan_extras = {"meval": True, "strict": False, "normalize": True}


# ---------------------------------------------------------------------- fused device path of the CLI loop
def native_chain(models, device, fp16, normalize=False):
    """The engines of a model chain if the whole per-image loop of run.py:421-434 can stay on the device
    (np2tensor -> [chop_forward per model] -> tensor2np [-> color_fix]), else None.  That needs a CUDA device, chop
    mode, and every model to be a 3-channel network with a native engine.  With ``normalize`` (images mapped to [-1, 1]
    around the chain, run.py:420,430) it needs a single CycleGAN generator, whose engine folds the mapping into its first
    and last layer (``unit_io``); every other case keeps the tensor interface."""
    if torch.device(device).type != "cuda":
        return None
    dtype = torch.float16 if fp16 else torch.float32
    engines = []
    for m in models:
        net = m.model
        if m.arch == "ts" or not m.chop or not hasattr(net, "native_engine") or m.in_nc != 3 or m.out_nc != 3:
            return None
        if normalize:
            from .architectures.ResNet_arch import ResnetGenerator
            if len(models) != 1 or not isinstance(net, ResnetGenerator):
                return None
            engines.append(net.native_engine(device, dtype, unit_io=True))
        else:
            engines.append(net.native_engine(device, dtype))
    return engines


class ChainRunner:
    """uint8 HWC BGR frame in, uint8 frame out, everything in between on the device: H2D of the frame from a pinned
    staging buffer, image -> tiles (np2tensor fused) -> model 1 -> fp16/fp32 tensor -> ... -> last model -> blend to
    uint8 (tensor2np fused) -> optional -cf colour fix (three kernels) -> one D2H into pinned memory."""

    def __init__(self, engines, device, cf=False, patch_size=200, step=0.5):
        self.engines, self.device, self.cf = engines, torch.device(device), cf
        self.patch_size, self.step = patch_size, step
        self._pin_in = self._pin_out = None

    def _pinned(self, attr, shape):
        buf = getattr(self, attr)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, dtype=torch.uint8).pin_memory()
            setattr(self, attr, buf)
        return buf

    def run_device(self, d_img):
        """Device uint8 [H,W,3] -> device uint8 [S*H,S*W,3] (no synchronisation)."""
        x = d_img
        for i, eng in enumerate(self.engines):
            x = eng.chop_forward_ex(x, self.patch_size, self.step, out_u8=(i == len(self.engines) - 1))
        if self.cf:
            x = color_fix_device(d_img, x)
        return x

    def __call__(self, img):
        """Host uint8 HWC BGR numpy image -> host uint8 numpy image (a view of the runner's pinned result buffer,
        valid until the next call)."""
        if img.dtype != "uint8" or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected a uint8 HWC 3-channel image")
        stage = self._pinned("_pin_in", img.shape)
        stage.numpy()[...] = img
        d_img = stage.to(self.device, non_blocking=True)
        d_out = self.run_device(d_img)
        res = self._pinned("_pin_out", d_out.shape)
        res.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return res.numpy()


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("-models", "-m", type=str, required=True, help="Path to models.")
    p.add_argument("-arch", "-a", type=str, required=False, default="infer", help="Model architecture.")
    p.add_argument("-input", "-i", type=str, required=False, default="./input", help="Path to read input images.")
    p.add_argument("-output", "-o", type=str, required=False, default="./output", help="Path to save output images.")
    p.add_argument("-scale", "-s", type=str, required=False, default="-1", help="Model scaling factor.")
    p.add_argument("-cf", required=False, action="store_true", help="Use color correction if enabled.")
    p.add_argument("-comp", required=False, action="store_true", help="Save as comparison images if enabled.")
    p.add_argument("-no_gpu", "-cpu", required=False, action="store_false", help="Run in CPU if enabled.")
    p.add_argument("-no_fp16", required=False, action="store_false", help="Disable fp16 mode if needed.")
    p.add_argument("-norm", required=False, action="store_true",
                   help="Normalizes images in range [-1,1] if set, else [0,1].")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    gpu = args.no_gpu                      # store_false flags: True means "use the GPU"
    fp16 = False if args.arch == "ts" else (args.no_fp16 and gpu)
    if "wbc" in args.arch or "wbc" in args.models:
        raise NotImplementedError("the WBC (white-box cartoonisation) family exists in the reference but is outside the "
                                  "scope of this engine (SURVEY.md section 8f)")
    extras, chop, resize = default_extras, True, False
    if "unet_" in args.arch or "p2p_" in args.arch:          # run.py:347-356
        extras, chop = pix2pix_extras, False
        resize = next((n for n in (512, 256, 128) if str(n) in args.arch), False)
    elif "resnet_" in args.arch or "cg_" in args.arch:       # run.py:357-360
        extras = cyclegan_extras
    meval, strict = extras["meval"], extras["strict"]
    normalize = extras["normalize"] or args.norm

    if gpu:
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available; pass -cpu to run the torch CPU path explicitly")
        device = torch.device("cuda")
    else:
        device = torch.device("cpu")

    model_chain, scale_chain = parse_models(args.models)
    models = []
    for path, scale in zip(model_chain, scale_chain):
        m = Model(path, args.arch, scale, device=device, meval=meval, strict=strict, chop=chop)
        if fp16 and args.arch != "ts":
            m.model.half()
        models.append(m)

    engines = native_chain(models, device, fp16, normalize=normalize)
    runner = ChainRunner(engines, device, cf=args.cf) if engines else None

    for image_path in get_images_paths(args.input):
        name = osp.splitext(osp.basename(image_path))[0]
        img = read_img(image_path)
        if img is None:
            print(f"Error reading image {image_path}, skipping.")
            continue
        if resize:
            img = linear_resize(img, resize)
        if runner is not None and img.dtype == "uint8" and img.ndim == 3 and img.shape[2] == 3:
            # 3-channel uint8 image through 3-channel native models: the whole loop body below as one device pipeline
            img_out = runner(img).copy()
            out_path = osp.join(args.output, f"{name:s}.png")
            if args.comp:
                save_img_comp([img, img_out], out_path)
            else:
                save_img(img_out, out_path)
            continue
        t_img = np2tensor(img, normalize=normalize).to(device)
        t_img = t_img.half() if fp16 else t_img
        t_out = t_img.clone()
        for mod in models:
            t_out = mod(t_out)
        img_out = tensor2np(t_out.detach(), denormalize=normalize)
        if args.cf:
            img_out = color_fix(img, img_out, device=device)
        out_path = osp.join(args.output, f"{name:s}.png")
        if args.comp:
            save_img_comp([img, img_out], out_path)
        else:
            save_img(img_out, out_path)


if __name__ == "__main__":
    main()
