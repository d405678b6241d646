"""GPU parity at the EXACT shapes BASELINE.json's configs put on the hot path (SURVEY.md
Appendix B), against the UNMODIFIED reference CUDA extension (oracle/_ref/GANet*.so) run on
the same device -- it needs about 1.3 s for a 920M-voxel sample, so even the headline
microbench sample is compared directly rather than through properties:

  config 2 (GANet-11, 240x624)      SGA 1x32x65x80x208 (M3), 1x48x33x40x104 (M6); LGA2 1x193x240x624
  config 3 (GANet-deep, 384x1248)   SGA 1x32x65x128x416 (K3), 1x48x33x64x208 (K6); LGA2 1x193x384x1248
  config 4 (training step)          the shapes of config 2, forward and backward
  config 5 (microbench sweep)       SGA 1x32x{96,192}x240x624 and 1x16x288x240x624 (the reference
                                    indexes with int: one call must stay below 2^31 elements),
                                    LGA2 1x{96,192,288}x240x624

Criteria: SGA forward values, direction mask and depth arg-max bit-exact;
gradients and LGA within 1e-4 relative fp32.  Both backward variants (recompute / kept
aggregates) are checked.  Everything is compared on the device.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_gpu

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_gpu.available(), reason="reference CUDA extension not built")]
RTOL = 1e-4          # north_star: 1e-4 relative fp32

SGA_CONFIG_SHAPES = {
    "M3": (1, 32, 65, 80, 208), "M6": (1, 48, 33, 40, 104),
    "K3": (1, 32, 65, 128, 416), "K6": (1, 48, 33, 64, 208),
    "sweep_D96": (1, 32, 96, 240, 624), "headline_D192": (1, 32, 192, 240, 624),
    "sweep_D288": (1, 16, 288, 240, 624),
}
LGA_CONFIG_SHAPES = {
    "L": (1, 193, 240, 624), "L_kitti": (1, 193, 384, 1248),
    "sweep_D96": (1, 96, 240, 624), "headline_D192": (1, 192, 240, 624), "sweep_D288": (1, 288, 240, 624),
}


def _need(nbytes):
    free, _ = torch.cuda.mem_get_info()
    if free < nbytes:
        pytest.skip("needs %.0f GB of free device memory" % (nbytes / 2 ** 30))


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _close(a, b, what):
    """SURVEY.md 8c, both halves, on the device: per-tensor max|a-b| / max|b| <= 1e-4 AND the
    element-wise allclose(rtol=1e-4, atol=1e-5 * scale)."""
    assert a.shape == b.shape, what
    scale = float(b.abs().max().clamp_min(1e-30))
    diff = (a - b).abs()
    err = float(diff.max()) / scale
    assert err <= RTOL, "%s: relative error %.3g > %.1g" % (what, err, RTOL)
    excess = diff - (1e-5 * scale + RTOL * b.abs())
    bad = int((excess > 0).sum())
    assert bad == 0, "%s: %d of %d elements outside allclose(rtol=1e-4, atol=1e-5*scale), worst excess %.3g" % (
        what, bad, a.numel(), float(excess.max()))


@pytest.fixture(scope="module")
def ops():
    from ganet_b200 import ops as o
    return o


@pytest.mark.parametrize("name", list(SGA_CONFIG_SHAPES))
def test_sga_config_shape_vs_reference_cuda(ops, name):
    shape = SGA_CONFIG_SHAPES[name]
    N, C, D, H, W = shape
    numel = N * C * D * H * W
    assert numel < 2 ** 31
    _need(70 * numel)                 # ~17 volumes live at the peak
    torch.manual_seed(11 + D)
    x = torch.randn(shape, device="cuda")
    g = [F.normalize(torch.randn(N, C, 5, H, W, device="cuda"), p=1, dim=2) for _ in range(4)]
    go = torch.randn(shape, device="cuda")

    ro, rm, rtemp = ref_gpu.sga_forward(x, *g)
    out, mask = ops.sga_forward(x, *g)
    assert torch.equal(out, ro), "SGA forward values are not bit-identical"
    assert torch.equal(mask, rm.to(torch.uint8)), "direction mask differs"
    del out, mask
    out2, mask2, agg = ops.sga_forward(x, *g, keep_aggregates=True)
    assert torch.equal(out2, ro) and torch.equal(mask2, rm.to(torch.uint8))
    del ro, out2

    rgi, rgg, ridx = ref_gpu.sga_backward(x, *g, rtemp, rm, go)
    del rtemp, rm
    gi, gg, idx = ops.sga_backward(x, *g, mask2, go, want_max_idx=True)
    assert torch.equal(idx, ridx.to(torch.int32)), "depth arg-max differs"
    _close(gi, rgi, "gradInput")
    for k in range(4):
        _close(gg[k], rgg[k], "guidance gradient %d" % k)
    gi2, gg2 = ops.sga_backward(x, *g, mask2, go, aggregates=agg)
    assert torch.equal(gi2, gi) and all(torch.equal(a, b) for a, b in zip(gg2, gg))


@pytest.mark.parametrize("name", list(LGA_CONFIG_SHAPES))
def test_lga2_config_shape_vs_reference_cuda(ops, name):
    shape = LGA_CONFIG_SHAPES[name]
    N, D, H, W = shape
    torch.manual_seed(5 + D)
    x = torch.randn(shape, device="cuda")
    f = F.normalize(torch.randn(N, 75, H, W, device="cuda"), p=1, dim=1)
    go = torch.randn(shape, device="cuda")
    ry, ry1 = ref_gpu.lga2_forward(x, f)
    rgx, rgf = ref_gpu.lga2_backward(x, f, ry1, go.clone())
    y1 = ops.lga_forward(x, f, 2)
    y = ops.lga_forward(y1, f, 2)
    _close(y1, ry1, "LGA pass 1")
    _close(y, ry, "LGA2 output")
    g1, gf = ops.lga_backward(y1, f, go, 2)
    gx, gf = ops.lga_backward(x, f, g1, 2, gf)
    _close(gx, rgx, "LGA2 grad_x")
    _close(gf, rgf, "LGA2 grad_filters")
