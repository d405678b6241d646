"""Import shim for libs/sync_bn/modules/sync_bn.py (reference :139-153).

The reference vendors PyTorch-Encoding's SyncBatchNorm; its CUDA extension does
not compile against torch 2.x and it is outside the hot
path.  The models only need the two class names.  Same constructor arguments,
parameters and state_dict keys (weight, bias, running_mean, running_var,
num_batches_tracked); under DistributedDataParallel convert with
torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).
"""
import torch.nn as nn


class BatchNorm1d(nn.BatchNorm1d):
    pass


class BatchNorm2d(nn.BatchNorm2d):
    pass


class BatchNorm3d(nn.BatchNorm3d):
    pass


SyncBatchNorm = nn.SyncBatchNorm

__all__ = ["BatchNorm1d", "BatchNorm2d", "BatchNorm3d", "SyncBatchNorm"]
