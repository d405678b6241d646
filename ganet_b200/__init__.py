"""ganet_b200 -- B200-native (sm_100a) guided-aggregation operators for GA-Net.

Hot path only: SGA, LGA, GetCostVolume, DisparityRegression behind the
reference's torch.autograd.Function / nn.Module names, over a C-ABI CUDA library
(include/ganet_b200.h, ganet_b200/csrc/).  See DESIGN.md.
"""
__version__ = "0.1.0"
