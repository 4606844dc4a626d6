"""ctypes loader for libewvit.so.  The signatures mirror include/ewvit.h one to one."""
import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

P = c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_LOCK = threading.Lock()
_LIB = None


class EwvitError(RuntimeError):
    """A libewvit.so entry point returned a negative status."""


def lib_path():
    return os.environ.get("EWVIT_LIB", os.path.join(_HERE, "libewvit.so"))


# name -> (restype, argtypes); must list EVERY symbol declared in include/ewvit.h
SIGNATURES = {
    "ewvit_abi_version": (c_int, []),
    "ewvit_last_error": (c_char_p, []),
    "ewvit_launch_count": (c_uint64, []),
    "ewvit_dwt_haar_fwd": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ewvit_dwt3_haar_fwd": (c_int, [c_void_p, c_int64, c_int, c_int] + [c_void_p] * 6 + [c_void_p]),
    "ewvit_linear_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "ewvit_conv3x3_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ewvit_maxpool2x2_nhwc_bf16": (c_int, [P, c_int64, c_int, c_int, c_int, P, P]),
    "ewvit_gap_nhwc_bf16": (c_int, [P, c_int64, c_int, c_int, P, c_int64, P]),
    "ewvit_vit_assemble": (c_int, [P, P, P, P, c_int64, c_int, c_int, P, P]),
    "ewvit_layernorm_bf16": (c_int, [P, c_int64, P, P, c_float, P, c_int64, c_int64, c_int, P]),
    "ewvit_vit_attention": (c_int, [P, c_int64, c_int, c_int, c_int, P, P]),
    "ewvit_dama_wpack_floats": (c_int64, [c_int, c_int]),
    "ewvit_dama_tail_fwd": (c_int, [P, P, c_int64, c_int, c_int, c_int, P, c_float, P, P, P, P]),
    "ewvit_conv_nhwc_bf16": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P, P]),
    "ewvit_conv3x3_c24_fwd": (c_int, [P, P, c_int, P, c_int, c_int, c_int, c_int, P, P]),
    "ewvit_dwt3_haar_u8_fwd": (c_int, [P, P, P, c_int, c_int64, c_int, c_int, P, P, P, P, P, P, P]),
    "ewvit_stem_conv_u8_fwd": (c_int, [P, P, P, c_int, c_int, c_int, P, P, c_int, P, c_int, P]),
    "ewvit_stem_conv_fwd": (c_int, [P, c_int, c_int, c_int, P, P, c_int, P, P]),
    "ewvit_stem_conv_padded_fwd": (c_int, [P, c_int, c_int, c_int, P, P, c_int, P, P]),
    "ewvit_stem_conv_same_fwd": (c_int, [P, c_int, c_int, c_int, P, P, c_int, P, P]),
    "ewvit_conv_nhwc_bf16_ex": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P, c_int, c_int, P]),
    "ewvit_dwconv3x3_nhwc_bf16": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "ewvit_se_gate_fwd": (c_int, [P, c_int, P, P, P, P, c_int, c_int, c_int, P, c_int, P]),
    "ewvit_dwconv_nhwc_bf16": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "ewvit_dwconv_pool_parts": (c_int, [c_int, c_int, c_int, c_int]),
    "ewvit_conv1x1_gated_nhwc_bf16": (c_int, [P, P, c_int, P, c_int, c_int, c_int, c_int, P, c_int, P, P, P]),
    "ewvit_mwt_upsample3_fwd": (c_int, [P, P, P, c_int, c_int, c_int, P, P]),
    "ewvit_mwt_head_conv3_fwd": (c_int, [P, P, c_int, c_int, c_int, P, P, P, P]),
    "ewvit_binary_metrics_fwd": (c_int, [P, P, c_int, P, P]),
    "ewvit_debug_set_trace": (c_int, [P]),
    "ewvit_debug_set_flags": (c_int, [c_int]),
    "ewvit_video_head_fwd": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, P, P, P, P, P, c_int, P, P]),
}


def load():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _LIB
    with _LOCK:
        if _LIB is None:
            path = lib_path()
            if not os.path.exists(path):
                raise EwvitError(
                    f"{path} not found: build it with `python efficient-wavelet-vit_b200/build.py` "
                    "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
            handle = ctypes.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)          # AttributeError if the .so is stale
                fn.restype, fn.argtypes = res, args
            _LIB = handle
    return _LIB


def lib():
    return load()


def check(status, what):
    if status != 0:
        msg = load().ewvit_last_error()
        raise EwvitError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")
