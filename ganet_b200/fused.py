"""Fusions around the hot-path operators inside the reference's MODELS, applied to a model
object without touching models/*.py.

`fuse_sga_blocks(model)` gives every `SGABlock` (models/GANet_deep.py:249-277, GANet11.py:231-259) a forward
whose prologue -- split of the guidance conv output into four direction blocks, view as (N, C, 5, H, W),
F.normalize(p=1, dim=2) of each -- is ONE kernel (functions.SgaGuidanceFunction) instead of a dozen torch
launches and their autograd nodes; the rest of the block (SGA, BatchNorm + ReLU, the 3-D refinement conv, the
residual) runs as the reference wrote it.  Values are bit-identical to the unfused block
(tests/test_gpu_fused.py), so SGA's max / arg-max decisions do not move.
"""
import types

import torch
import torch.nn.functional as F

from .functions import NormDispRegressionFunction, SgaGuidanceFunction


def _fused_sga_block_forward(self, x, g):
    rem = x
    k1, k2, k3, k4 = SgaGuidanceFunction.apply(g.contiguous(), x.size()[1])      # :264-268
    x = self.SGA(x, k1, k2, k3, k4)                                             # :269
    if self.refine:                                                             # :270-274
        x = self.bn_relu(x)
        x = self.conv_refine(x)
    else:
        x = self.bn(x)
    assert x.size() == rem.size()
    x += rem
    return self.relu(x)


def fuse_sga_blocks(model):
    """Patch every SGABlock instance of `model` (matched by class name and attributes, so that the reference's
    files need no import from here); returns the number of blocks fused."""
    n = 0
    for m in model.modules():
        if type(m).__name__ == "SGABlock" and hasattr(m, "SGA") and hasattr(m, "refine"):
            m.forward = types.MethodType(_fused_sga_block_forward, m)
            n += 1
    return n


def unfuse_sga_blocks(model):
    n = 0
    for m in model.modules():
        if type(m).__name__ == "SGABlock" and "forward" in m.__dict__:
            del m.forward
            n += 1
    return n


def _fused_disp_agg_forward(self, x, lg1, lg2):
    """DispAgg.forward (models/GANet_deep.py:239-247) with its last two steps -- F.normalize(p=1, dim=1) and the
    disparity regression -- as one kernel each way (partial)."""
    x = F.interpolate(self.conv32x1(x), [self.maxdisp + 1, x.size()[3] * 3, x.size()[4] * 3], mode='trilinear',
                      align_corners=False)
    x = torch.squeeze(x, 1)
    assert lg1.size() == lg2.size()
    x = self.lga(x, lg1)
    x = self.softmax(x)
    x = self.lga(x, lg2)
    return NormDispRegressionFunction.apply(x.contiguous())


def fuse_disp_heads(model):
    n = 0
    for m in model.modules():
        if type(m).__name__ == "DispAgg" and hasattr(m, "LGA2") and hasattr(m, "conv32x1"):
            m.forward = types.MethodType(_fused_disp_agg_forward, m)
            n += 1
    return n


def unfuse_disp_heads(model):
    n = 0
    for m in model.modules():
        if type(m).__name__ == "DispAgg" and "forward" in m.__dict__:
            del m.forward
            n += 1
    return n
