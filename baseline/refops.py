"""Harness only: run a model's hot-path modules on the UNMODIFIED reference CUDA extension
(oracle/_ref/GANet*.so, built by oracle/build_ref.py) instead of ganet_b200's kernels, so the same
model object with the same weights gives the reference's answer and the reference's timing on the
same GPU.  The autograd wrappers restate libs/GANet/functions/GANet.py (SgaFunction :8-48,
Lga2Function :174-203) over the extension's own entry points with its buffer contract
(caller-zeroed outputs); GetCostVolume / DisparityRegression are the reference's lines
(modules/GANet.py:119-148) as written.  Nothing under ganet_b200/ imports this file."""
import contextlib

import torch
from torch.autograd import Function

from oracle import ref_gpu


class RefSga(Function):
    @staticmethod
    def forward(ctx, x, g0, g1, g2, g3):
        out, mask, temp = ref_gpu.sga_forward(x, g0, g1, g2, g3)
        ctx.save_for_backward(x, g0, g1, g2, g3, temp, mask)
        return out

    @staticmethod
    def backward(ctx, go):
        x, g0, g1, g2, g3, temp, mask = ctx.saved_tensors
        gi, gg, _ = ref_gpu.sga_backward(x, g0, g1, g2, g3, temp, mask, go.contiguous())
        return (gi,) + tuple(gg)


class RefLga2(Function):
    @staticmethod
    def forward(ctx, x, f):
        y, y1 = ref_gpu.lga2_forward(x, f, 2)
        ctx.save_for_backward(x, f, y1)
        return y

    @staticmethod
    def backward(ctx, go):
        x, f, y1 = ctx.saved_tensors
        gx, gf = ref_gpu.lga2_backward(x, f, y1, go.contiguous().clone(), 2)
        return gx, gf


def ref_cost_volume(x, y, dm):                     # modules/GANet.py:119-134 as written
    num, channels, height, width = x.size()
    cost = x.new_zeros(num, channels * 2, dm, height, width)
    for i in range(dm):
        if i > 0:
            cost[:, :channels, i, :, i:] = x[:, :, :, i:]
            cost[:, channels:, i, :, i:] = y[:, :, :, :-i]
        else:
            cost[:, :channels, i] = x
            cost[:, channels:, i] = y
    return cost.contiguous()


def ref_disp_regression(p):                        # modules/GANet.py:142-148 as written
    disp = torch.arange(p.shape[1], device=p.device, dtype=p.dtype).reshape(1, -1, 1, 1)
    return torch.sum(p * disp.repeat(p.size(0), 1, p.size(2), p.size(3)), 1)


def hot_path_modules(model):
    """(name, module) of every hot-path operator instance the model calls."""
    import ganet_b200.modules as M
    kinds = (M.SGA, M.LGA2, M.GetCostVolume, M.DisparityRegression)
    return [(n, m) for n, m in model.named_modules() if isinstance(m, kinds)]


@contextlib.contextmanager
def reference_ops(model):
    """Inside the block the model's SGA / LGA2 / GetCostVolume / DisparityRegression instances
    call the reference implementation (instance-level forward override; the model is not rebuilt)."""
    import ganet_b200.modules as M
    if not ref_gpu.available():
        raise RuntimeError("oracle/_ref/GANet*.so not built")
    patched = []
    for _, m in hot_path_modules(model):
        if isinstance(m, M.SGA):
            m.forward = lambda x, g0, g1, g2, g3: RefSga.apply(x.contiguous(), g0.contiguous(), g1.contiguous(),
                                                               g2.contiguous(), g3.contiguous())
        elif isinstance(m, M.LGA2):
            m.forward = lambda x, f: RefLga2.apply(x.contiguous(), f.contiguous())
        elif isinstance(m, M.GetCostVolume):
            m.forward = (lambda mod: (lambda x, y: ref_cost_volume(x, y, mod.maxdisp)))(m)
        else:
            m.forward = ref_disp_regression
        patched.append(m)
    try:
        yield model
    finally:
        for m in patched:
            del m.forward          # back to the class's forward
