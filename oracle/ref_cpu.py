"""numpy front-end over oracle/_ref/libganet_ref_cpu.so -- the reference's own
kernel bodies compiled for the host (see oracle/build_ref.py).
TEST INFRASTRUCTURE ONLY.

The buffer contract is the reference's: the caller zero-fills outputs and
scratch (libs/GANet/functions/GANet.py:14-16, :33-39), masks are float volumes.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libganet_ref_cpu.so")
_lib = None
_f = ctypes.POINTER(ctypes.c_float)


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libganet_ref_cpu.so missing: run "
                               "`python oracle/build_ref.py --cpu-only` where /root/reference exists")
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(_f)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sga_forward(x, g0, g1, g2, g3):
    """SgaFunction.forward (functions/GANet.py:10-22) -> out, mask(f32), temp_out (left aggregate)"""
    x, g0, g1, g2, g3 = (_c(a) for a in (x, g0, g1, g2, g3))
    N, C, D, H, W = x.shape
    out = np.zeros_like(x); temp = np.zeros_like(x); mask = np.zeros_like(x)
    rc = lib().ref_sga_forward(_p(x), _p(g0), _p(g1), _p(g2), _p(g3), _p(temp), _p(out), _p(mask),
                               N, C, D, H, W)
    assert rc == 1, rc
    return out, mask, temp


def sga_backward(x, g0, g1, g2, g3, temp_out, mask, grad_out):
    """SgaFunction.backward (functions/GANet.py:24-48) -> grad_in, (g0..g3 grads), max_idx (f32)"""
    x, g0, g1, g2, g3, mask, grad_out = (_c(a) for a in (x, g0, g1, g2, g3, mask, grad_out))
    temp_out = _c(temp_out).copy()          # the reference overwrites it (:1064)
    N, C, D, H, W = x.shape
    gi = np.zeros_like(x); tg = np.zeros_like(x)
    gg = [np.zeros_like(g0) for _ in range(4)]
    idx = np.zeros((N, C, H, W), np.float32)
    rc = lib().ref_sga_backward(_p(x), _p(g0), _p(g1), _p(g2), _p(g3), _p(temp_out), _p(mask),
                                _p(idx), _p(grad_out), _p(tg), _p(gi), _p(gg[0]), _p(gg[1]),
                                _p(gg[2]), _p(gg[3]), N, C, D, H, W)
    assert rc == 1, rc
    return gi, tuple(gg), idx


def _dims(x):
    lead = int(np.prod(x.shape[:-3]))
    return (lead,) + tuple(x.shape[-3:])


def lga_forward(x, f, radius=2):
    x, f = _c(x), _c(f)
    B, D, H, W = _dims(x)
    y = np.zeros_like(x)
    rc = lib().ref_lga_forward(_p(x), _p(f), _p(y), B, D, H, W, radius)
    assert rc == 1
    return y


def lga_backward(x, f, grad_out, grad_f=None, radius=2):
    """one lga_backward call; grad_f accumulates if given"""
    x, f, grad_out = _c(x), _c(f), _c(grad_out)
    B, D, H, W = _dims(x)
    gx = np.zeros_like(x)
    gf = np.zeros_like(f) if grad_f is None else grad_f
    rc = lib().ref_lga_backward(_p(x), _p(f), _p(grad_out), _p(gx), _p(gf), B, D, H, W, radius)
    assert rc == 1
    return gx, gf


def lga2_forward(x, f, radius=2):
    """Lga2Function.forward (functions/GANet.py:176-187) -> y, y1"""
    y1 = lga_forward(x, f, radius)
    return lga_forward(y1, f, radius), y1


def lga2_backward(x, f, y1, grad_out, radius=2):
    """Lga2Function.backward (functions/GANet.py:189-203) -> gx, gf"""
    g1, gf = lga_backward(y1, f, grad_out, None, radius)
    gx, gf = lga_backward(x, f, g1, gf, radius)
    return gx, gf
