"""Import shim for libs/sync_bn/modules/sync_bn.py (reference :26-153).

The reference vendors PyTorch-Encoding's SyncBatchNorm; its CUDA extension does
not compile against torch 2.x and it is outside the hot
path.  The models only need the class names (models/GANet_deep.py:8).

Constructor: the reference's `SyncBatchNorm(num_features, eps=1e-5, momentum=0.1,
sync=True, activation="none", slope=0.01, inplace=True)` (:64-66).  `sync`,
`slope` and `inplace` are accepted and ignored; an `activation` other than "none"
is applied after the normalisation ("leaky_relu" with `slope`, "relu").  Parameters
and state_dict keys are torch's (`weight, bias, running_mean, running_var,
num_batches_tracked`), the reference's own (:27,81).

Statistics: plain per-process batch statistics.  Cross-GPU statistics -- what the
reference's master/worker queue does under nn.DataParallel (:86-113) -- come from
one process per GPU: `torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)` turns
every class below into torch's NCCL-synchronised layer, then wrap the model in
DistributedDataParallel (tests/test_gpu_ddp.py, bench.py --config 4).  Under
nn.DataParallel with more than one replica the statistics stay per replica; a
warning says so once.
"""
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

_warned = False


class _ShimMixin:
    def _shim_init(self, sync, activation, slope, inplace):
        if activation not in ("none", "leaky_relu", "relu"):
            raise ValueError("sync_bn shim: unknown activation %r" % (activation,))
        self.sync, self.activation, self.slope, self.inplace = sync, activation, slope, inplace

    def forward(self, x):
        global _warned
        if (self.training and not _warned and x.is_cuda and torch.cuda.device_count() > 1
                and not torch.distributed.is_initialized()
                and torch.cuda.current_device() != 0):
            # a replica other than device 0 is running outside torch.distributed: DataParallel
            _warned = True
            warnings.warn("libs.sync_bn shim: batch statistics are per replica under nn.DataParallel; "
                          "use one process per GPU with nn.SyncBatchNorm.convert_sync_batchnorm + "
                          "DistributedDataParallel for synchronised statistics")
        y = super().forward(x)
        if self.activation == "leaky_relu":
            return F.leaky_relu(y, self.slope)
        if self.activation == "relu":
            return F.relu(y)
        return y


class BatchNorm1d(_ShimMixin, nn.BatchNorm1d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, sync=True, activation="none", slope=0.01,
                 inplace=True):
        nn.BatchNorm1d.__init__(self, num_features, eps=eps, momentum=momentum, affine=True)
        self._shim_init(sync, activation, slope, inplace)


class BatchNorm2d(_ShimMixin, nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, sync=True, activation="none", slope=0.01,
                 inplace=True):
        nn.BatchNorm2d.__init__(self, num_features, eps=eps, momentum=momentum, affine=True)
        self._shim_init(sync, activation, slope, inplace)


class BatchNorm3d(_ShimMixin, nn.BatchNorm3d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, sync=True, activation="none", slope=0.01,
                 inplace=True):
        nn.BatchNorm3d.__init__(self, num_features, eps=eps, momentum=momentum, affine=True)
        self._shim_init(sync, activation, slope, inplace)


SyncBatchNorm = nn.SyncBatchNorm

__all__ = ["BatchNorm1d", "BatchNorm2d", "BatchNorm3d", "SyncBatchNorm"]
