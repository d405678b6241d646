"""nn.Module layer -- host-side mirror of libs/GANet/modules/GANet.py.

Same class names, constructor arguments and forward signatures as the reference,
so models/GANet_deep.py and models/GANet11.py run unchanged on these modules
(see libs/ at the repository root, which re-exports them under the reference's
import paths).  None of them owns parameters or buffers, exactly as upstream, so
state_dict keys of the models are unaffected.
"""
import torch
from torch.nn.modules.module import Module

from .functions import (CostVolumeFunction, DisparityRegressionFunction, Lga2Function,
                        Lga3d2Function, Lga3d3Function, Lga3dFunction, Lga3Function,
                        LgaFunction, MyLoss2Function, MyLossFunction, SgaFunction)


class MyNormalize(Module):
    """MyNormalize (modules/GANet.py:18-33): signed L1 normalisation along `dim`
    with a +-1e-6 guard.  Imported by the models, never instantiated there."""

    def __init__(self, dim):
        self.dim = dim
        super(MyNormalize, self).__init__()

    def forward(self, x):
        norm = torch.sum(torch.abs(x), self.dim, keepdim=True)
        # upstream applies two masked updates in sequence (<= 0: -1e-6, then
        # >= 0: +1e-6); sums of |x| are >= 0, so zeros end at -1e-6 and every
        # positive sum is shifted by +1e-6
        norm = torch.where(norm > 0, norm + 1e-6, norm - 1e-6)
        return torch.div(x, norm)


class MyLoss2(Module):
    """MyLoss2 (modules/GANet.py:34-41), used by train.py:71 for KITTI fine-tuning."""

    def __init__(self, thresh=1, alpha=2):
        super(MyLoss2, self).__init__()
        self.thresh = thresh
        self.alpha = alpha

    def forward(self, input1, input2):
        return MyLoss2Function.apply(input1, input2, self.thresh, self.alpha)


class MyLoss(Module):
    """MyLoss (modules/GANet.py:42-49); like upstream the constructor arguments
    are ignored and the thresholds are fixed at 5 / 1."""

    def __init__(self, upper_thresh=5, lower_thresh=1):
        super(MyLoss, self).__init__()
        self.upper_thresh = 5
        self.lower_thresh = 1

    def forward(self, input1, input2):
        return MyLossFunction.apply(input1, input2, self.upper_thresh, self.lower_thresh)


class SGA(Module):
    """SGA (modules/GANet.py:52-58): forward(input, g0, g1, g2, g3) with
    input (N,C,D,H,W) and L1-normalised guidance (N,C,5,H,W) per direction."""

    def __init__(self):
        super(SGA, self).__init__()

    def forward(self, input, g0, g1, g2, g3):
        return SgaFunction.apply(input, g0, g1, g2, g3)


class _LGABase(Module):
    _fn = None

    def __init__(self, radius=2):
        super(_LGABase, self).__init__()
        self.radius = radius

    def forward(self, input1, input2):
        return self._fn.apply(input1, input2, self.radius)


class LGA3D3(_LGABase):
    """LGA3D3 (modules/GANet.py:62-69)."""
    _fn = Lga3d3Function


class LGA3D2(_LGABase):
    """LGA3D2 (modules/GANet.py:70-77)."""
    _fn = Lga3d2Function


class LGA3D(_LGABase):
    """LGA3D (modules/GANet.py:78-85)."""
    _fn = Lga3dFunction


class LGA3(_LGABase):
    """LGA3 (modules/GANet.py:87-94)."""
    _fn = Lga3Function


class LGA2(_LGABase):
    """LGA2 (modules/GANet.py:95-102): the variant both models call
    (models/GANet_deep.py:236)."""
    _fn = Lga2Function


class LGA(_LGABase):
    """LGA (modules/GANet.py:103-110)."""
    _fn = LgaFunction


class GetCostVolume(Module):
    """GetCostVolume (modules/GANet.py:114-134): (N,C,H,W) x 2 -> (N,2C,maxdisp+1,H,W)."""

    def __init__(self, maxdisp):
        super(GetCostVolume, self).__init__()
        self.maxdisp = maxdisp + 1

    def forward(self, x, y):
        assert x.is_contiguous() == True  # noqa: E712  (upstream's precondition)
        return CostVolumeFunction.apply(x, y.contiguous(), self.maxdisp)


class DisparityRegression(Module):
    """DisparityRegression (modules/GANet.py:136-148): soft arg-min expectation
    sum_d d * p[:, d] over maxdisp+1 planes."""

    def __init__(self, maxdisp):
        super(DisparityRegression, self).__init__()
        self.maxdisp = maxdisp + 1

    def forward(self, x):
        assert x.is_contiguous() == True  # noqa: E712
        if x.shape[1] != self.maxdisp:
            # upstream: the (1, maxdisp, 1, 1) index tensor would not broadcast
            raise RuntimeError("DisparityRegression: expected %d disparity planes, got %d"
                               % (self.maxdisp, x.shape[1]))
        return DisparityRegressionFunction.apply(x)
