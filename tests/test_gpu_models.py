"""BASELINE.json configs 2 and 3 on the GPU box: the reference's OWN models -- models/GANet11.py and
models/GANet_deep.py, copied unmodified into baseline/_ref/models/ at build time -- run on the new
operators through the drop-in `libs.GANet` / `libs.sync_bn` import surface, and the same model
object with the same weights is run again with its hot-path modules switched to the UNMODIFIED
reference CUDA extension (baseline/refops.py).

  * every SGA call inside the model is checked bit-for-bit against the reference extension on the
    very tensors the model feeds it (forward values bit-exact);
  * every LGA2 / GetCostVolume / DisparityRegression call likewise, <= 1e-4 (exact for the copy);
  * the aggregated cost volume that enters the disparity head (models/GANet_deep.py:359) is
    bit-identical between the two runs -- everything before the head is SGA, GetCostVolume and cuDNN;
  * the disparity maps of the two runs (models/GANet_deep.py:389-410, models/GANet11.py:311-353)
    agree: the head is LGA2 -> softmin -> LGA2 -> F.normalize(p=1) -> regression, and with freshly
    initialised (signed) LGA filters its last two steps divide a cancelling sum by sum|x|, so a 1e-6
    difference after LGA2 (allowed: 1e-4) is amplified at ill-conditioned pixels.  Criterion: median
    error <= 1e-4 of the disparity range and >= 99 % of the pixels within 1e-3; the distribution
    (median, 99 %, worst pixel) is printed.

BatchNorm statistics: a freshly initialised model in eval mode has running_mean 0 / var 1, which lets
activations grow through ~50 conv layers until the soft-argmin saturates; like a trained checkpoint
would, the running statistics are therefore first set from one training-mode forward of the same
input pair (momentum 1), then both runs use eval mode.
"""
import numpy as np
import pytest
import torch

from baseline import refmodels, refops
from oracle import ref_gpu
from util import assert_close

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_gpu.available(), reason="reference CUDA extension not built"),
              pytest.mark.skipif(not refmodels.available(), reason="reference models not copied")]


def _calibrate_bn(model, left, right):
    bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    old = [m.momentum for m in bns]
    for m in bns:
        m.momentum = 1.0
    model.train()
    with torch.no_grad():
        model(left, right)
    for m, mo in zip(bns, old):
        m.momentum = mo
    model.eval()


def _check_calls_against_reference(model, log):
    """Forward hooks on the model's hot-path modules: each call's inputs go through the reference
    implementation as well and the outputs are compared on the spot."""
    import ganet_b200.modules as M
    handles = []

    def hook(mod, inputs, output):
        inputs = [t.contiguous() for t in inputs]
        if isinstance(mod, M.SGA):
            ref = ref_gpu.sga_forward(*inputs)[0]
            assert torch.equal(output, ref), "SGA output differs from the reference extension"
            log.append(("SGA", tuple(output.shape)))
        elif isinstance(mod, M.LGA2):
            ref = ref_gpu.lga2_forward(*inputs)[0]
            assert_close(output.cpu().numpy(), ref.cpu().numpy(), what="LGA2 inside the model")
            log.append(("LGA2", tuple(output.shape)))
        elif isinstance(mod, M.GetCostVolume):
            assert torch.equal(output, refops.ref_cost_volume(inputs[0], inputs[1], mod.maxdisp))
            log.append(("GetCostVolume", tuple(output.shape)))
        else:
            assert_close(output.cpu().numpy(), refops.ref_disp_regression(inputs[0]).cpu().numpy(),
                         what="DisparityRegression inside the model")
            log.append(("DisparityRegression", tuple(output.shape)))

    for _, m in refops.hot_path_modules(model):
        handles.append(m.register_forward_hook(hook))
    return handles


@pytest.mark.parametrize("name,H,W,n_sga", [("GANet11", 240, 624, 4), ("GANet_deep", 384, 1248, 7)])
def test_reference_model_inference_on_new_operators(name, H, W, n_sga):
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True      # the two runs must be comparable bit for bit
    dev = torch.device("cuda:0")
    model = refmodels.build(name, 192, seed=0, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    left = torch.randn(1, 3, H, W, device=dev, generator=gen)
    right = torch.randn(1, 3, H, W, device=dev, generator=gen)
    _calibrate_bn(model, left, right)

    log = []
    handles = _check_calls_against_reference(model, log)
    head = [m for m in model.modules() if type(m).__name__ == "DispAgg"]
    assert len(head) == 1
    seen = []
    handles.append(head[0].register_forward_pre_hook(lambda mod, inp: seen.append(inp[0].clone())))
    with torch.no_grad():
        d_new = model(left, right)
    for h in handles[:-1]:
        h.remove()
    assert d_new.shape == (1, H, W)
    assert [k for k, _ in log].count("SGA") == n_sga
    assert [k for k, _ in log].count("LGA2") == 2
    assert [k for k, _ in log].count("GetCostVolume") == 1
    assert [k for k, _ in log].count("DisparityRegression") == 1
    sga_shapes = sorted({s for k, s in log if k == "SGA"})
    assert sga_shapes == sorted({(1, 32, 65, H // 3, W // 3), (1, 48, 33, H // 6, W // 6)})

    with refops.reference_ops(model), torch.no_grad():
        d_ref = model(left, right)
    handles[-1].remove()
    assert len(seen) == 2 and torch.equal(seen[0], seen[1]), \
        "the aggregated cost volume entering the disparity head differs between the two runs"
    a, b = d_new.cpu().numpy().astype(np.float64), d_ref.cpu().numpy().astype(np.float64)
    assert np.isfinite(b).all() and b.std() > 1e-3, "degenerate disparity map"
    err = np.abs(a - b) / 192.0                       # relative to the disparity range (max_disp)
    within = float((err <= 1e-3).mean())
    print("\n%s %dx%d disparity, new operators vs reference extension: median %.3g, 99%% %.3g, max %.3g of the "
          "disparity range; %.4f of the pixels within 1e-3" % (name, H, W, np.median(err),
                                                             np.quantile(err, 0.99), err.max(), within))
    assert np.median(err) <= 1e-4
    assert within >= 0.99


def test_reference_model_training_step_on_new_operators():
    """GANet-11 in training mode on a 96x192 crop (multiples of 48, README.md:63): the train.py loss
    (:116-118, SceneFlow branch), the three disparity maps and the parameter gradients, new
    operators vs reference extension from the same weights.  The loss must agree to 1e-4; the
    disparity maps are compared like the inference test does (the heads amplify rounding-level
    differences of their inputs at ill-conditioned pixels); the parameter gradients pass through
    ~40 layers of cuDNN convolutions and atomically accumulated interpolation gradients after the
    operators, so they are compared per tensor at 1e-2 of the tensor's scale and, all together, by
    the cosine of the two gradient vectors (>= 0.9999)."""
    import torch.nn.functional as F
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    dev = torch.device("cuda:0")
    H, W = 96, 192
    model = refmodels.build("GANet11", 192, seed=3, device=dev).train()
    gen = torch.Generator(device=dev).manual_seed(4)
    left = torch.randn(1, 3, H, W, device=dev, generator=gen)
    right = torch.randn(1, 3, H, W, device=dev, generator=gen)
    target = torch.rand(1, H, W, device=dev, generator=gen) * 191.0

    def step():
        model.zero_grad(set_to_none=True)
        disp1, disp2 = model(left, right)
        loss = 0.4 * F.smooth_l1_loss(disp1, target) + 1.2 * F.smooth_l1_loss(disp2, target)
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        return loss.item(), disp1.detach().cpu().numpy(), disp2.detach().cpu().numpy(), grads

    l_new, d1_new, d2_new, g_new = step()
    with refops.reference_ops(model):
        l_ref, d1_ref, d2_ref, g_ref = step()
    assert abs(l_new - l_ref) <= 1e-4 * abs(l_ref)
    for what, a, b in (("disp1", d1_new, d1_ref), ("disp2", d2_new, d2_ref)):
        err = np.abs(a.astype(np.float64) - b) / 192.0
        print("\ntraining-mode %s: median %.3g, 99%% %.3g, max %.3g of the disparity range"
              % (what, np.median(err), np.quantile(err, 0.99), err.max()))
        assert np.median(err) <= 1e-4 and (err <= 1e-3).mean() >= 0.99, what
    assert set(g_new) == set(g_ref) and len(g_new) >= 180
    worst, worst_name, dot, na, nb = 0.0, None, 0.0, 0.0, 0.0
    for n in g_ref:
        a, b = g_new[n].double(), g_ref[n].double()
        scale = max(float(b.abs().max()), 1e-30)
        e = float((a - b).abs().max()) / scale
        if e > worst:
            worst, worst_name = e, n
        dot += float((a * b).sum()); na += float((a * a).sum()); nb += float((b * b).sum())
    cos = dot / max((na * nb) ** 0.5, 1e-300)
    print("parameter gradients: worst per-tensor relative error %.3g (%s), cosine %.8f" % (worst, worst_name, cos))
    assert cos >= 0.9999
    assert worst <= 1e-2, "parameter gradients differ: worst relative error %.3g (%s)" % (worst, worst_name)
