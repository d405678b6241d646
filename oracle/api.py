"""numpy front-end over oracle/libganet_oracle.so.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libganet_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_u8 = ctypes.POINTER(ctypes.c_uint8)
_i32 = ctypes.POINTER(ctypes.c_int32)
_i64 = ctypes.c_int64


def build(force=False):
    src = os.path.join(_HERE, "ganet_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libganet_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t=_f):
    return None if a is None else a.ctypes.data_as(t)


def _c(a, dt=np.float32):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


def num_threads():
    return int(lib().oracle_num_threads())


def sga_forward(x, g0, g1, g2, g3, fused=True, want_dirs=False):
    """-> out (N,C,D,H,W) f32, mask (N,C,D,H,W) u8 [, dirs (4,N,C,D,H,W)]"""
    x, g0, g1, g2, g3 = (_c(a) for a in (x, g0, g1, g2, g3))
    N, C, D, H, W = x.shape
    assert g0.shape == (N, C, 5, H, W)
    out = np.empty_like(x)
    mask = np.empty(x.shape, np.uint8)
    dirs = np.empty((4,) + x.shape, np.float32) if want_dirs else None
    rc = lib().oracle_sga_forward(_p(x), _p(g0), _p(g1), _p(g2), _p(g3), _p(out), _p(mask, _u8),
                                  _p(dirs), _i64(N), _i64(C), _i64(D), _i64(H), _i64(W),
                                  ctypes.c_int(int(fused)))
    assert rc == 0
    return (out, mask, dirs) if want_dirs else (out, mask)


def sga_backward(x, g0, g1, g2, g3, mask, grad_out, fused=True):
    """-> grad_in, (gg0, gg1, gg2, gg3), max_idx (N,C,H,W) int32 of the right scan"""
    x, g0, g1, g2, g3, grad_out = (_c(a) for a in (x, g0, g1, g2, g3, grad_out))
    mask = _c(mask, np.uint8)
    N, C, D, H, W = x.shape
    gi = np.empty_like(x)
    gg = [np.empty_like(g0) for _ in range(4)]
    idx = np.empty((N, C, H, W), np.int32)
    rc = lib().oracle_sga_backward(_p(x), _p(g0), _p(g1), _p(g2), _p(g3), _p(mask, _u8),
                                   _p(grad_out), _p(gi), _p(gg[0]), _p(gg[1]), _p(gg[2]),
                                   _p(gg[3]), _p(idx, _i32), _i64(N), _i64(C), _i64(D),
                                   _i64(H), _i64(W), ctypes.c_int(int(fused)))
    assert rc == 0
    return gi, tuple(gg), idx


def _lga_dims(x, f, radius):
    lead = int(np.prod(x.shape[:-3]))
    D, H, W = x.shape[-3:]
    F = 3 * (2 * radius + 1) ** 2
    assert f.shape == x.shape[:-3] + (F, H, W), (f.shape, x.shape)
    return lead, D, H, W


def lga_forward(x, f, radius=2, passes=1):
    """x: (N,D,H,W) or (N,C,D,H,W); f: (N,75,H,W) or (N,C,75,H,W).
    -> y, intermediates (passes-1, ...)"""
    x, f = _c(x), _c(f)
    B, D, H, W = _lga_dims(x, f, radius)
    y = np.empty_like(x)
    tmp = np.empty((max(passes - 1, 1),) + x.shape, np.float32)
    rc = lib().oracle_lga_multi_forward(_p(x), _p(f), _p(y), _p(tmp), _i64(B), _i64(D), _i64(H),
                                        _i64(W), ctypes.c_int(radius), ctypes.c_int(passes))
    assert rc == 0
    return y, tmp[:passes - 1]


def lga_backward(x, f, tmp, grad_out, radius=2, passes=1):
    """-> grad_x, grad_f"""
    x, f, grad_out = _c(x), _c(f), _c(grad_out)
    tmp = _c(tmp) if passes > 1 else np.zeros((1,) + x.shape, np.float32)
    B, D, H, W = _lga_dims(x, f, radius)
    gx = np.empty_like(x)
    gf = np.empty_like(f)
    scratch = np.empty((2,) + x.shape, np.float32)
    rc = lib().oracle_lga_multi_backward(_p(x), _p(f), _p(tmp), _p(grad_out), _p(gx), _p(gf),
                                         _p(scratch), _i64(B), _i64(D), _i64(H), _i64(W),
                                         ctypes.c_int(radius), ctypes.c_int(passes))
    assert rc == 0
    return gx, gf


def cost_volume_forward(x, y, maxdisp_plus1):
    x, y = _c(x), _c(y)
    N, C, H, W = x.shape
    Dm = int(maxdisp_plus1)
    cost = np.empty((N, 2 * C, Dm, H, W), np.float32)
    lib().oracle_cost_volume_forward(_p(x), _p(y), _p(cost), _i64(N), _i64(C), _i64(Dm),
                                     _i64(H), _i64(W))
    return cost


def cost_volume_backward(gcost):
    gcost = _c(gcost)
    N, C2, Dm, H, W = gcost.shape
    C = C2 // 2
    gx = np.empty((N, C, H, W), np.float32)
    gy = np.empty((N, C, H, W), np.float32)
    lib().oracle_cost_volume_backward(_p(gcost), _p(gx), _p(gy), _i64(N), _i64(C), _i64(Dm),
                                      _i64(H), _i64(W))
    return gx, gy


def disp_regression_forward(p):
    p = _c(p)
    N, Dm, H, W = p.shape
    out = np.empty((N, H, W), np.float32)
    lib().oracle_disp_regression_forward(_p(p), _p(out), _i64(N), _i64(Dm), _i64(H), _i64(W))
    return out


def disp_regression_backward(gdisp, Dm):
    gdisp = _c(gdisp)
    N, H, W = gdisp.shape
    gp = np.empty((N, Dm, H, W), np.float32)
    lib().oracle_disp_regression_backward(_p(gdisp), _p(gp), _i64(N), _i64(Dm), _i64(H), _i64(W))
    return gp
