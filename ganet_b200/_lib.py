"""ctypes binding of libganet_b200.so (the C ABI in include/ganet_b200.h).

The product path has NO fallback: if the CUDA library is missing or a tensor is
not a CUDA fp32 tensor this module raises; nothing here ever routes to a CPU
or PyTorch implementation.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libganet_b200.so")

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ganet_b200.h one to one
_SIGNATURES = {
    "ganet_abi_version": (_int, []),
    "ganet_error_string": (ctypes.c_char_p, [_int]),
    "ganet_sga_forward": (_int, [_vp] * 9 + [_sz] + [_i64] * 5 + [_vp]),
    "ganet_sga_forward_workspace_min": (_sz, [_i64] * 5),
    "ganet_sga_forward_workspace_best": (_sz, [_i64] * 5),
    "ganet_sga_backward_workspace_min": (_sz, [_i64] * 5),
    "ganet_sga_backward_workspace_best": (_sz, [_i64] * 5),
    "ganet_sga_aggregate_volumes": (_int, [_i64] * 5),
    "ganet_sga_backward": (_int, [_vp] * 15 + [_sz] + [_i64] * 5 + [_vp]),
    "ganet_sga_direction": (_int, [_vp] * 3 + [_int] + [_i64] * 5 + [_vp]),
    "ganet_lga_forward": (_int, [_vp] * 3 + [_i64] * 4 + [_int, _vp]),
    "ganet_lga_backward": (_int, [_vp] * 5 + [_int] + [_i64] * 4 + [_int, _vp]),
    "ganet_cost_volume_forward": (_int, [_vp] * 3 + [_i64] * 5 + [_vp]),
    "ganet_cost_volume_backward": (_int, [_vp] * 3 + [_i64] * 5 + [_vp]),
    "ganet_disp_regression_forward": (_int, [_vp] * 2 + [_i64] * 4 + [_vp]),
    "ganet_disp_regression_backward": (_int, [_vp] * 2 + [_i64] * 4 + [_vp]),
    "ganet_norm_disp_regression_forward": (_int, [_vp] * 3 + [_i64] * 4 + [_vp]),
    "ganet_norm_disp_regression_backward": (_int, [_vp] * 5 + [_i64] * 4 + [_vp]),
    "ganet_sga_guidance_forward": (_int, [_vp] * 5 + [_i64] * 4 + [_vp]),
    "ganet_sga_guidance_backward": (_int, [_vp] * 6 + [_i64] * 4 + [_vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class GanetNativeError(RuntimeError):
    pass


def lib():
    """Load the native library once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise GanetNativeError(
                "ganet_b200: native CUDA library not built (%s missing). "
                "Run `python -m ganet_b200.build` (needs nvcc); there is no fallback path." % SO_PATH)
        handle = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.ganet_abi_version() != 2:
            raise GanetNativeError("ganet_b200: ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().ganet_error_string(rc).decode()
        raise GanetNativeError("ganet_b200 native call failed (%d): %s" % (rc, msg))


def stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor of the expected dtype."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GanetNativeError("ganet_b200 ops run on CUDA tensors only (got %s); there is no CPU path"
                               % t.device)
    if t.dtype != dtype:
        raise TypeError("ganet_b200: expected %s tensor, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise GanetNativeError("ganet_b200: tensor must be contiguous")
    return _vp(t.data_ptr())
