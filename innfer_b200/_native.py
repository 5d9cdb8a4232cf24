"""ctypes binding of the C-ABI library (include/innfer_b200.h).

There is no CPU fallback behind these calls: if the shared library is missing or a call fails,
an exception is raised.  Use ``innfer_b200.build.build()`` (or ``__graft_entry__.build()``) to
compile the library in-tree.
"""
import ctypes
import os

from . import build as _build

INNFER_F16, INNFER_F32, INNFER_U8 = 0, 1, 2


class NativeError(RuntimeError):
    """A call into libinnfer_b200.so returned a negative status."""

    def __init__(self, code, message):
        super().__init__("innfer_b200 native error %d: %s" % (code, message))
        self.code = code


class RRDBCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("in_nc", "out_nc", "nf", "nb", "gc", "scale", "plus", "fp16")]


class SRResNetCfg(ctypes.Structure):
    _fields_ = [("in_nc", ctypes.c_int32), ("out_nc", ctypes.c_int32), ("nf", ctypes.c_int32), ("nb", ctypes.c_int32),
                ("scale", ctypes.c_int32), ("upsample_mode", ctypes.c_int32), ("res_scale", ctypes.c_float),
                ("fp16", ctypes.c_int32)]


class PPONCfg(ctypes.Structure):
    _fields_ = [("in_nc", ctypes.c_int32), ("out_nc", ctypes.c_int32), ("nf", ctypes.c_int32), ("nb", ctypes.c_int32),
                ("scale", ctypes.c_int32), ("alpha", ctypes.c_float), ("fp16", ctypes.c_int32)]


class PANCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("in_nc", "out_nc", "nf", "unf", "nb", "scale", "self_attention", "double_scpa", "fp16")]


class I2ICfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("kind", "in_nc", "out_nc", "ngf", "depth", "norm", "train", "fp16", "unit_io")]


class Tile(ctypes.Structure):
    _fields_ = [("y0", ctypes.c_int32), ("x0", ctypes.c_int32)]


_LIB = None

_vp, _i, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double
_u64, _u32 = ctypes.c_uint64, ctypes.c_uint32
_PROTOS = {
    "innfer_last_error": (ctypes.c_char_p, []),
    "innfer_version": (ctypes.c_char_p, []),
    "innfer_kernel_launches": (ctypes.c_uint64, []),
    "innfer_rrdb_create": (_i, [ctypes.POINTER(RRDBCfg), _i, ctypes.POINTER(_vp)]),
    "innfer_srresnet_create": (_i, [ctypes.POINTER(SRResNetCfg), _i, ctypes.POINTER(_vp)]),
    "innfer_ppon_create": (_i, [ctypes.POINTER(PPONCfg), _i, ctypes.POINTER(_vp)]),
    "innfer_pan_create": (_i, [ctypes.POINTER(PANCfg), _i, ctypes.POINTER(_vp)]),
    "innfer_i2i_create": (_i, [ctypes.POINTER(I2ICfg), _i, ctypes.POINTER(_vp)]),
    "innfer_rrdb_load": (_i, [_vp, ctypes.c_char_p, _vp, ctypes.POINTER(ctypes.c_int64), _i]),
    "innfer_rrdb_finalize": (_i, [_vp]),
    "innfer_rrdb_destroy": (None, [_vp]),
    "innfer_rrdb_set_max_batch": (_i, [_vp, _i]),
    "innfer_rrdb_profile": (_i, [_vp, _i]),
    "innfer_rrdb_profile_read": (_i, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "innfer_rrdb_profile_families": (_i, [_vp, ctypes.c_char_p, _u64, ctypes.POINTER(_u64)]),
    "innfer_rrdb_chop_forward_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _i, _vp]),
    "innfer_rrdb_tile_bytes": (_i, [_vp, _i, _i, _i, _d, ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(_i)]),
    "innfer_device_memset": (_i, [_vp, _i, _u64]),
    "innfer_memcpy_async": (_i, [_vp, _vp, _u64, _vp]),
    "innfer_stream_signal": (_i, [ctypes.POINTER(_vp), _i, _u32, _vp]),
    "innfer_stream_wait": (_i, [ctypes.POINTER(_vp), _i, _u32, _vp, _u64, _vp]),
    "innfer_blend_f32": (_i, [_vp, _i, _i, _i, _d, _i, _i, _vp, _i, _vp]),
    "innfer_rrdb_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "innfer_rrdb_chop_forward": (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _i, _vp]),
    "innfer_rrdb_upscale_u8": (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _vp]),
    "innfer_rrdb_upscale_u8_device": (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _vp]),
    "innfer_rrdb_tile_buffer": (_i, [_vp, _i, _i, _i, _d, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_uint64),
                                     ctypes.POINTER(ctypes.c_uint64)]),
    "innfer_rrdb_forward_tile_range": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _i, _i, _vp, _vp]),
    "innfer_rrdb_blend_tiles": (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _i, _vp]),
    "innfer_ipc_export": (_i, [_vp, _vp]),
    "innfer_ipc_open": (_i, [_vp, ctypes.POINTER(_vp)]),
    "innfer_ipc_close": (_i, [_vp]),
    "innfer_device_alloc": (_i, [_i, ctypes.c_uint64, ctypes.POINTER(_vp)]),
    "innfer_device_free": (_i, [_vp]),
    "innfer_debug_set_trace": (_i, [_vp]),
    "innfer_debug_conv_loop": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_float)]),
    "innfer_device_upload": (_i, [_vp, _vp, ctypes.c_uint64, _vp]),
    "innfer_tiles_plan": (_i, [_i, _i, _i, _d, ctypes.POINTER(Tile), _i, ctypes.POINTER(_i),
                               ctypes.POINTER(_i)]),
    "innfer_image_to_tiles": (_i, [_vp, _i, _i, _i, _i, _i, _d, _vp, _vp]),
    "innfer_blend": (_i, [_vp, _i, _i, _i, _d, _i, _i, _vp, _i, _vp]),
    "innfer_conv3x3": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _f, _vp, _i, _i, _vp]),
    "innfer_gen_conv": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp]),
    "innfer_debug_i2i_halo_launches": (ctypes.c_uint64, []),
    "innfer_debug_i2i_graph_replays": (ctypes.c_uint64, []),
    "innfer_color_fix": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    "innfer_color_fix_host": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i]),
}

EXPORTED_SYMBOLS = tuple(sorted(_PROTOS))


def lib_path():
    return _build.lib_path()


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "innfer_b200: %s is missing. Build it with `python -m innfer_b200.build` "
            "(needs nvcc). There is no CPU fallback for the CUDA path." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def last_error():
    return load().innfer_last_error().decode(errors="replace")


def check(rc):
    if rc != 0:
        raise NativeError(rc, last_error())


def kernel_launches():
    return int(load().innfer_kernel_launches())
