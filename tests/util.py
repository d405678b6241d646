"""Shared input generators for the tests (numpy side and torch side)."""
import numpy as np


def l1norm(a, axis):
    return (a / np.abs(a).sum(axis=axis, keepdims=True)).astype(np.float32)


def sga_inputs(shape, seed=0, positive=False):
    """x, [g0..g3] L1-normalised over dim 2 (what SGABlock feeds, models/GANet_deep.py:265),
    grad_out -- all float32 numpy."""
    rng = np.random.default_rng(seed)
    N, C, D, H, W = shape
    x = rng.standard_normal(shape).astype(np.float32)
    g = [rng.standard_normal((N, C, 5, H, W)) for _ in range(4)]
    if positive:      # x in [1, 1.1), weights > 0: every aggregate stays in [1, 1.1)
        x = (1.0 + 0.1 * rng.random(shape)).astype(np.float32)
        g = [np.abs(a) + 0.05 for a in g]
    g = [l1norm(a, 2) for a in g]
    go = rng.standard_normal(shape).astype(np.float32)
    return x, g, go


def lga_inputs(shape, seed=0, radius=2):
    rng = np.random.default_rng(seed)
    F = 3 * (2 * radius + 1) ** 2
    x = rng.standard_normal(shape).astype(np.float32)
    fs = tuple(shape[:-3]) + (F,) + tuple(shape[-2:])
    f = l1norm(rng.standard_normal(fs), len(fs) - 3)
    go = rng.standard_normal(shape).astype(np.float32)
    return x, f, go


def rel_err(a, b):
    """max|a-b| / max(max|b|, tiny) -- the per-tensor relative error of SURVEY.md 8c"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def assert_close(a, b, rtol=1e-4, what=""):
    """north_star tolerance: 1e-4 relative fp32.  Per-tensor relative error plus an
    element-wise allclose whose atol is scaled to the tensor's magnitude."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(float(np.abs(b).max()), 1e-30)
    err = rel_err(a, b)
    assert err <= rtol, "%s: relative error %.3g > %.1g" % (what, err, rtol)
    assert np.allclose(a, b, rtol=rtol, atol=rtol * scale), what
