"""ctypes binding of libstb200.so (the C ABI declared in include/stb200.h).

There is no fallback: if the shared library is missing or a kernel reports an error the call
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (or
``stereo_toolbox_b200/csrc/build.sh``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STB200_LIB") or os.path.join(_HERE, "libstb200.so")     # STB200_LIB: A/B builds of the same source (kernel work)

_lib = None


class StbError(RuntimeError):
    pass


_P, _I, _F, _LL = c_void_p, c_int, c_float, c_longlong
_IP = POINTER(c_int)
_PP = POINTER(c_void_p)

# name -> argtypes ; every function returns int except the two library-level ones
SIGNATURES = {
    "stb_gwc_volume_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_concat_volume_f32": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_sampled_volume_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_softmax_d_f32": [_P, _P, _I, _I, _LL, _P],
    "stb_upsample_softargmin_f32": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_disparity_regression_f32": [_P, _P, _I, _I, _LL, _P],
    "stb_conv3d_taps_f32": [_P, _P, _P, _P, _P] + [_I] * 10 + [_IP, _IP, _IP] + [_I] * 9 + [_P],
    "stb_volume_cl16": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_volume_cl16_from_cl16": [_PP, _IP, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_ncdhw_to_cl16": [_P, _P, _I, _I, _I, _LL, _I, _P],
    "stb_cl16_to_ncdhw": [_P, _P, _I, _I, _I, _LL, _I, _P],
    "stb_conv3d_umma": [_P, _P, _P, _P, _P, _P] + [_I] * 13 + [_IP, _IP, _IP, _IP, _IP, _IP, _IP, _I, _I, _IP, _IP, _IP, _IP, _IP]
                       + [_I] * 11 + [_P],
    "stb_conv3d_umma_set_trace": [_P, _I],
    "stb_conv3d_taps_cl16": [_P, _P, _P, _P, _P, _I, _I] + [_I] * 10 + [_IP, _IP, _IP] + [_I] * 9 + [_P],
    "stb_corr1d_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "stb_avgpool_last_f32": [_P, _P, _LL, _I, _P],
    "stb_corr1d_lookup_f32": [_PP, _P, _LL, _P, _I, _I, _I, _I, _I, _I, _P],
    "stb_geo_lookup_f32": [_PP, _PP, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_geo_permute_f32": [_P, _P, _I, _I, _I, _I, _I, _P],
    "stb_patch_dw_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_block_attention": [_P, _P, _P] + [_I] * 10 + [POINTER(c_longlong), POINTER(c_longlong), _P],
    "stb_feature_gate_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "stb_feature_gate_cl16": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "stb_disparity_variance_f32": [_P, _P, _P, _I, _I, _LL, _P],
    "stb_convex_upsample_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "stb_context_upsample_f32": [_P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "stb_conv3d_wgrad_f32": [_P, _P, _P] + [_I] * 12 + [_P],
    "stb_conv3d_wgrad_cl16": [_P, _P, _P] + [_I] * 14 + [_P],
    "stb_warp_disp_f32": [_P, _P, _P, _I, _I, _I, _I, _P],
    "stb_corr_volume_1d_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "stb_gru_rh_split": [_P, _P, _P, _LL, _I, _P],
    "stb_gru_blend_split": [_P, _P, _P, _P, _LL, _I, _P],
    "stb_pool2x_split": [_P, _P, _I, _I, _I, _I, _P],
    "stb_interp_split": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "stb_concat_volume_bwd_f32": [_P, _P, _P] + [_I] * 8 + [_P],
    "stb_gwc_volume_bwd_f32": [_P, _P, _P, _P, _P] + [_I] * 8 + [_P],
    "stb_upsample_softargmin_bwd_f32": [_P, _P, _P] + [_I] * 8 + [_P],
}


def lib():
    """Load (once) and return the ctypes handle; raises StbError if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StbError(
                f"{LIB_PATH} not found: the CUDA library is not built and there is no CPU fallback. "
                "Run stereo_toolbox_b200/csrc/build.sh (or __graft_entry__.build()).")
        h = ctypes.CDLL(LIB_PATH)
        h.stb_error_string.restype = c_char_p
        h.stb_error_string.argtypes = [c_int]
        h.stb_version.restype = c_int
        for name, argtypes in SIGNATURES.items():
            fn = getattr(h, name)      # AttributeError here = header/library mismatch
            fn.restype = c_int
            fn.argtypes = argtypes
        _lib = h
    return _lib


def register(name, argtypes):
    """Late registration used by optional translation units (e.g. the tcgen05 conv path)."""
    SIGNATURES[name] = argtypes
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes


def check(code: int, what: str):
    if code != 0:
        raise StbError(f"{what} failed: {lib().stb_error_string(code).decode()} (code {code})")


LAUNCH_COUNT = 0   # number of kernel launches issued through this binding (bench.py reports it)

# Device of the tensors of the call being assembled.  The C ABI launches on the CURRENT device; the wrappers build a call's
# arguments with ops._p(tensor) / ops._stream() and then invoke call(): _p notes the device of every tensor it sees (and
# refuses a mix), _stream() returns that device's current stream, and call() makes that device current for the launch --
# so ``model.to('cuda:1')`` works without torch.cuda.set_device, like the device-guarded torch ops of the reference.
_call_device = None


def note_device(index):
    """Called by ops._p for every CUDA tensor argument of the call being assembled."""
    global _call_device
    if _call_device is None:
        _call_device = index
    elif _call_device != index:
        bad, _reset = (_call_device, index), _reset_device()
        raise StbError(f"tensor arguments of one stereo_toolbox_b200 call live on different CUDA devices {bad}")


def _reset_device():
    global _call_device
    _call_device = None


def call_device():
    return _call_device


def call(name: str, *args):
    global LAUNCH_COUNT
    dev = _call_device
    _reset_device()
    fn = getattr(lib(), name)
    if dev is None:
        code = fn(*args)
    else:
        import torch
        if torch.cuda.current_device() == dev:
            code = fn(*args)
        else:
            with torch.cuda.device(dev):
                code = fn(*args)
    check(code, name)
    LAUNCH_COUNT += 1
