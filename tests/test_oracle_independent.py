"""A second, independent pin of the oracle: SURVEY.md Appendix A (the semantics of the hot path
as read off the reference kernels, GANet_kernel.cu:66-933, :1131-1269) restated here in float64
numpy with explicit loops -- written from the formulas, sharing no code with oracle/ganet_oracle.c
-- must agree with the C oracle on small volumes.  (The oracle is additionally pinned bit-exactly
against the reference's own kernel bodies in test_oracle.py; this file guards against an error
common to both restatements of the launch sequences.)
"""
import numpy as np
import pytest

from oracle import api
from util import lga_inputs, sga_inputs


def _scan_positions(direction, H, W):
    """Appendix A.1 table: for each lane the list of (h, w) in scan order."""
    if direction == 0:
        return [[(h, w) for h in range(H)] for w in range(W)]
    if direction == 1:
        return [[(H - 1 - h, w) for h in range(H)] for w in range(W)]
    if direction == 2:
        return [[(h, w) for w in range(W)] for h in range(H)]
    return [[(h, W - 1 - w) for w in range(W)] for h in range(H)]


def np_direction(x, g, direction):
    """A.1: one directional aggregation of one (n,c) slice; x (D,H,W), g (5,H,W) -> A (D,H,W)."""
    D, H, W = x.shape
    A = np.zeros_like(x)
    for line in _scan_positions(direction, H, W):
        P = None
        for t, (h, w) in enumerate(line):
            xv, wt = x[:, h, w], g[:, h, w]
            if t == 0:
                cur = xv * wt[0] + xv * wt[1] + xv * wt[2] + xv * wt[3] + xv * wt[4]
            else:
                up = np.concatenate(([xv[0]], P[:-1]))       # P[d-1], x[d] at d = 0
                dn = np.concatenate((P[1:], [xv[-1]]))       # P[d+1], x[d] at d = D-1
                cur = xv * wt[0] + P * wt[1] + up * wt[2] + dn * wt[3] + P.max() * wt[4]
            A[:, h, w] = cur
            P = cur
    return A


def np_backward_direction(x, g, A, T0, direction):
    """A.3 for one (n,c) slice: T0 = gradOut * [mask == k].  -> gI (D,H,W), gw (5,H,W), idx (H,W)"""
    D, H, W = x.shape
    gI = np.zeros_like(x)
    gw = np.zeros_like(g)
    idx = np.zeros((H, W), np.int64)
    for line in _scan_positions(direction, H, W):
        n = len(line)
        T = [T0[:, h, w].copy() for (h, w) in line]
        ks = [int(np.argmax(A[:, h, w])) for (h, w) in line]      # first maximum
        for t in range(n - 1, -1, -1):
            h, w = line[t]
            wt = g[:, h, w]
            if t + 1 < n:
                hn, wn = line[t + 1]
                wn_ = g[:, hn, wn]
                Tn = T[t + 1]
                T[t] += Tn * wn_[1]
                T[t][:-1] += Tn[1:] * wn_[2]                     # [d+1 < D] T[d+1, t+1] w2
                T[t][1:] += Tn[:-1] * wn_[3]                     # [d >= 1]  T[d-1, t+1] w3
            gI[:, h, w] += T[t] * wt[0]
            if t + 1 < n:
                s = (T[t + 1] * wn_[4]).sum()                    # uses the FINAL T[., t+1]
                T[t][ks[t]] += s
                gI[ks[t], h, w] += s * wt[0]
        for t in range(n):
            h, w = line[t]
            wt = g[:, h, w]
            gI[0, h, w] += T[t][0] * wt[2]
            gI[D - 1, h, w] += T[t][D - 1] * wt[3]
            idx[h, w] = ks[t]
            xv = x[:, h, w]
            gw[0, h, w] = (T[t] * xv).sum()
            if t >= 1:
                hp, wp = line[t - 1]
                Ap = A[:, hp, wp]
                gw[1, h, w] = (T[t] * Ap).sum()
                gw[2, h, w] = T[t][0] * xv[0] + (T[t][1:] * Ap[:-1]).sum()
                gw[3, h, w] = T[t][D - 1] * xv[D - 1] + (T[t][:-1] * Ap[1:]).sum()
                gw[4, h, w] = T[t].sum() * Ap[ks[t - 1]]
    return gI, gw, idx


@pytest.mark.parametrize("shape", [(1, 2, 5, 4, 6), (2, 1, 3, 5, 3), (1, 1, 1, 3, 4), (1, 1, 7, 1, 5)])
def test_sga_oracle_agrees_with_appendix_a_in_float64(shape):
    x, g, go = sga_inputs(shape, seed=100 + sum(shape))
    N, C, D, H, W = shape
    out, mask, dirs = api.sga_forward(x, *g, fused=False, want_dirs=True)
    gi, gg, idx = api.sga_backward(x, *g, mask, go, fused=False)
    x64, go64 = x.astype(np.float64), go.astype(np.float64)
    for n in range(N):
        for c in range(C):
            A = [np_direction(x64[n, c], g[k][n, c].astype(np.float64), k) for k in range(4)]
            for k in range(4):
                assert np.allclose(A[k], dirs[k][n, c], rtol=1e-5, atol=1e-6)
            best, win = A[0].copy(), np.zeros(A[0].shape, np.uint8)
            for k in (1, 2, 3):                                  # A.2: strict <, ties keep the lower id
                m = best < A[k]
                best[m] = A[k][m]
                win[m] = k
            assert np.allclose(best, out[n, c], rtol=1e-5, atol=1e-6)
            top2 = np.sort(np.stack(A), axis=0)[-2:]
            clear = (top2[1] - top2[0]) > 1e-5                   # away from fp32 near-ties
            assert np.array_equal(win[clear], mask[n, c][clear])
            # A.3 with the ORACLE's mask and fp32 aggregates (so that arg-max indices agree)
            gI = np.zeros((D, H, W))
            for k in (3, 0, 1, 2):
                T0 = go64[n, c] * (mask[n, c] == k)
                gIk, gwk, idxk = np_backward_direction(x64[n, c], g[k][n, c].astype(np.float64),
                                                       dirs[k][n, c].astype(np.float64), T0, k)
                gI += gIk
                assert np.allclose(gwk, gg[k][n, c], rtol=1e-4, atol=1e-5), k
                if k == 2:                                       # the reference leaves the last one
                    assert np.array_equal(idxk, idx[n, c])
            assert np.allclose(gI, gi[n, c], rtol=1e-4, atol=1e-5)


def np_lga(x, f, R):
    """A.4: y[d,h,w] = sum over taps f[loc,h,w] * x[d',h',w'], the centre voxel when out of range."""
    D, H, W = x.shape
    ws = 2 * R + 1
    y = np.zeros_like(x)
    for d in range(D):
        for h in range(H):
            for w in range(W):
                acc = 0.0
                for dd in (-1, 0, 1):
                    for r in range(-R, R + 1):
                        for c in range(-R, R + 1):
                            loc = (dd + 1) * ws * ws + (r + R) * ws + (c + R)
                            d2, h2, w2 = d + dd, h + r, w + c
                            inside = 0 <= d2 < D and 0 <= h2 < H and 0 <= w2 < W
                            acc += f[loc, h, w] * (x[d2, h2, w2] if inside else x[d, h, w])
                y[d, h, w] = acc
    return y


@pytest.mark.parametrize("radius", [2, 1, 0])
def test_lga_oracle_agrees_with_appendix_a_in_float64(radius):
    x, f, go = lga_inputs((2, 4, 5, 6), seed=3, radius=radius)
    y, _ = api.lga_forward(x, f, radius, 1)
    for n in range(2):
        ref = np_lga(x[n].astype(np.float64), f[n].astype(np.float64), radius)
        assert np.allclose(ref, y[n], rtol=1e-5, atol=1e-6)
