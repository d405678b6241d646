"""torch.autograd.Function layer -- host-side mirror of the reference's
libs/GANet/functions/GANet.py (same class names, argument order and returned
gradients), re-implemented over the C ABI.

Differences from the reference, all deliberate:
  * SgaFunction saves a 1-byte direction mask instead of two fp32 volumes
    (functions/GANet.py:21) and, when memory allows (ops.keep_aggregates_policy), the four
    directional aggregates, which lets backward skip its recompute passes;
  * backward never writes into its incoming gradOutput storage or into saved
    tensors (the reference does both, functions/GANet.py:197-200,
    GANet_kernel.cu:1064), so retain_graph / double use of a gradient works;
  * kernels run on the current stream, not legacy stream 0;
  * LgaFunction / Lga3Function, which are broken upstream (undefined `radius`,
    functions/GANet.py:241; typo `fitlers`, :155), work here with the signature
    the modules call them with;
  * Lgf2Function calls native entry points that never existed upstream
    (`lgf_cuda_*`, :216); it raises NotImplementedError here;
  * MyLoss2Function / MyLossFunction return None as the gradient of the target (input2);
    upstream returns a one-element CPU zero tensor (`Variable(torch.Tensor([0]))`, :289, :310),
    which current autograd rejects as soon as the target requires grad -- train.py never asks for it.
"""
import torch
from torch.autograd import Function

from . import ops


def _assert_contiguous(*tensors):
    # the reference's only precondition check (functions/GANet.py:11, :28, ...)
    for t in tensors:
        assert t.is_contiguous() == True  # noqa: E712  (same failure mode as upstream)


class SgaFunction(Function):
    """SgaFunction (functions/GANet.py:8-48): apply(input, g0, g1, g2, g3)."""

    @staticmethod
    def forward(ctx, input, g0, g1, g2, g3):
        _assert_contiguous(input, g0, g1, g2, g3)
        needs_bwd = any(t.requires_grad for t in (input, g0, g1, g2, g3))
        agg = None
        if ops.keep_aggregates_policy(input, needs_bwd):
            # memory-for-bandwidth: keep the four directional aggregates (16 B/voxel) so that
            # backward skips its recompute passes; ops.keep_aggregates_policy decides, and an
            # allocation failure falls back to the recompute variant instead of failing the step
            try:
                output, mask, agg = ops.sga_forward(input, g0, g1, g2, g3, keep_aggregates=True)
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
                agg = None
        if agg is not None:
            ctx.save_for_backward(input, g0, g1, g2, g3, mask, agg)
            ops.track_kept_aggregates(agg)           # counted until the buffer is freed, with or without a backward
        else:
            output, mask = ops.sga_forward(input, g0, g1, g2, g3)
            ctx.save_for_backward(input, g0, g1, g2, g3, mask)
        return output

    @staticmethod
    def backward(ctx, gradOutput):
        saved = ctx.saved_tensors
        input, g0, g1, g2, g3, mask = saved[:6]
        agg = saved[6] if len(saved) > 6 else None
        gradOutput = gradOutput.contiguous()
        gradInput, (grad0, grad1, grad2, grad3) = ops.sga_backward(input, g0, g1, g2, g3, mask,
                                                                   gradOutput, aggregates=agg)
        return gradInput, grad0, grad1, grad2, grad3


class SgaGuidanceFunction(Function):
    """SGABlock prologue, no reference counterpart as a Function: apply(g, channels) returns
    the four (N, C, 5, H, W) L1-normalised guidance tensors that models/GANet_deep.py:264-268 builds with
    torch.split + .view + F.normalize(p=1, dim=2) -- one kernel instead of a dozen, bit-identical values --
    and backward folds the four guidance gradients into the gradient of the raw guidance in one pass."""

    @staticmethod
    def forward(ctx, g, channels):
        _assert_contiguous(g)
        ctx.save_for_backward(g)
        return ops.sga_guidance_forward(g, channels)

    @staticmethod
    def backward(ctx, gg0, gg1, gg2, gg3):
        g, = ctx.saved_tensors
        return ops.sga_guidance_backward(g, gg0.contiguous(), gg1.contiguous(), gg2.contiguous(),
                                         gg3.contiguous()), None


class _LgaNFunction(Function):
    """`passes` successive LGA passes with shared filters; 4-D or 5-D input."""
    passes = 1

    @classmethod
    def _fwd(cls, ctx, input, filters, radius):
        ctx.radius = radius
        _assert_contiguous(input, filters)
        inter = []
        cur = input
        for _ in range(cls.passes):
            cur = ops.lga_forward(cur, filters, radius)
            inter.append(cur)
        ctx.save_for_backward(input, filters, *inter[:-1])
        return inter[-1]

    @classmethod
    def _bwd(cls, ctx, gradOutput):
        saved = ctx.saved_tensors
        input, filters, inter = saved[0], saved[1], list(saved[2:])
        g = gradOutput.contiguous()
        gradFilters = None
        for p in range(cls.passes - 1, -1, -1):
            src = input if p == 0 else inter[p - 1]
            g, gradFilters = ops.lga_backward(src, filters, g, ctx.radius, gradFilters)
        return g, gradFilters, None


class LgaFunction(_LgaNFunction):
    """LgaFunction (functions/GANet.py:239-263), one pass."""
    passes = 1

    @staticmethod
    def forward(ctx, input, filters, radius=1):
        return LgaFunction._fwd(ctx, input, filters, radius)

    @staticmethod
    def backward(ctx, gradOutput):
        return LgaFunction._bwd(ctx, gradOutput)


class Lga2Function(_LgaNFunction):
    """Lga2Function (functions/GANet.py:174-203), two passes; the one the models use."""
    passes = 2

    @staticmethod
    def forward(ctx, input, filters, radius=1):
        return Lga2Function._fwd(ctx, input, filters, radius)

    @staticmethod
    def backward(ctx, gradOutput):
        return Lga2Function._bwd(ctx, gradOutput)


class Lga3Function(_LgaNFunction):
    """Lga3Function (functions/GANet.py:141-173), three passes."""
    passes = 3

    @staticmethod
    def forward(ctx, input, filters, radius=1):
        return Lga3Function._fwd(ctx, input, filters, radius)

    @staticmethod
    def backward(ctx, gradOutput):
        return Lga3Function._bwd(ctx, gradOutput)


class Lga3dFunction(_LgaNFunction):
    """Lga3dFunction (functions/GANet.py:116-139): 5-D input, per-channel filters."""
    passes = 1

    @staticmethod
    def forward(ctx, input, filters, radius=1):
        return Lga3dFunction._fwd(ctx, input, filters, radius)

    @staticmethod
    def backward(ctx, gradOutput):
        return Lga3dFunction._bwd(ctx, gradOutput)


class Lga3d2Function(_LgaNFunction):
    """Lga3d2Function (functions/GANet.py:84-114)."""
    passes = 2

    @staticmethod
    def forward(ctx, input, filters, radius=1):
        return Lga3d2Function._fwd(ctx, input, filters, radius)

    @staticmethod
    def backward(ctx, gradOutput):
        return Lga3d2Function._bwd(ctx, gradOutput)


class Lga3d3Function(_LgaNFunction):
    """Lga3d3Function (functions/GANet.py:51-83)."""
    passes = 3

    @staticmethod
    def forward(ctx, input, filters, radius=1):
        return Lga3d3Function._fwd(ctx, input, filters, radius)

    @staticmethod
    def backward(ctx, gradOutput):
        return Lga3d3Function._bwd(ctx, gradOutput)


class Lgf2Function(Function):
    """Present upstream (functions/GANet.py:205-237) but bound to native symbols
    `lgf_cuda_forward/backward` that the reference never defined."""

    @staticmethod
    def forward(ctx, input, filters, radius=2):
        raise NotImplementedError("Lgf2Function: the reference has no lgf_cuda_* kernels either")

    @staticmethod
    def backward(ctx, gradOutput):
        raise NotImplementedError


class CostVolumeFunction(Function):
    """Autograd for GetCostVolume (modules/GANet.py:119-134): one kernel each way
    instead of 2*(maxdisp+1) CopySlices nodes."""

    @staticmethod
    def forward(ctx, x, y, maxdisp_plus1):
        return ops.cost_volume_forward(x, y, maxdisp_plus1)

    @staticmethod
    def backward(ctx, grad_cost):
        gx, gy = ops.cost_volume_backward(grad_cost.contiguous())
        return gx, gy, None


class DisparityRegressionFunction(Function):
    """Autograd for DisparityRegression (modules/GANet.py:142-148)."""

    @staticmethod
    def forward(ctx, x):
        ctx.dm = x.shape[1]
        return ops.disp_regression_forward(x)

    @staticmethod
    def backward(ctx, grad_disp):
        return ops.disp_regression_backward(grad_disp.contiguous(), ctx.dm)


class NormDispRegressionFunction(Function):
    """DispAgg tail (partial): F.normalize(x, p=1, dim=1) followed by DisparityRegression
    (models/GANet_deep.py:245-247) as one pass over x each way."""

    @staticmethod
    def forward(ctx, x):
        _assert_contiguous(x)
        disp, norm = ops.norm_disp_regression_forward(x)
        ctx.save_for_backward(x, disp, norm)
        return disp

    @staticmethod
    def backward(ctx, grad_disp):
        x, disp, norm = ctx.saved_tensors
        return ops.norm_disp_regression_backward(x, disp, norm, grad_disp.contiguous())


class MyLoss2Function(Function):
    """MyLoss2Function (functions/GANet.py:264-289).  Same piecewise values and
    gradient as upstream, including the order-dependent masked updates (each
    mask is evaluated on the partially updated tensor), without mutating the
    saved tensor."""

    @staticmethod
    def forward(ctx, input1, input2, thresh=1, alpha=2):
        ctx.thresh, ctx.alpha = thresh, alpha
        diff = input1 - input2
        ctx.save_for_backward(diff)
        v = diff.abs()
        v = torch.where(v < thresh, v * v / thresh, v)
        mid = (v <= thresh + alpha) & (v >= thresh)
        v = torch.where(mid, v * 2 - (v - thresh) ** 2 / (2.0 * alpha) - thresh, v)
        v = torch.where(v > thresh + alpha, v + alpha / 2.0, v)
        return v.mean()

    @staticmethod
    def backward(ctx, gradOutput):
        diff, = ctx.saved_tensors
        th, al = ctx.thresh, ctx.alpha
        s = diff.abs()
        s = torch.where(s > th + al, torch.ones_like(s), s)
        mid = (s <= th + al) & (s >= th)
        s = torch.where(mid, 2 - (s - th) / al, s)
        s = torch.where(s < th, 2 * s / th, s)
        grad = torch.sign(diff) * s * gradOutput / s.numel()
        return grad, None, None, None


class MyLossFunction(Function):
    """MyLossFunction (functions/GANet.py:291-310)."""

    @staticmethod
    def forward(ctx, input1, input2, upper_thresh=5, lower_thresh=1):
        ctx.upper_thresh, ctx.lower_thresh = upper_thresh, lower_thresh
        diff = input1 - input2
        ctx.save_for_backward(diff)
        return diff.abs().mean()

    @staticmethod
    def backward(ctx, gradOutput):
        diff, = ctx.saved_tensors
        up, lo = ctx.upper_thresh, ctx.lower_thresh
        s = diff.abs()
        s = torch.where(s > up, torch.ones_like(s), s)
        mid = (s <= up) & (s >= lo)
        s = torch.where(mid, 2 - (s - (up + lo) / 2.0).abs() / 2.0, s)
        grad = torch.sign(diff) * s * gradOutput
        return grad, None, None, None
