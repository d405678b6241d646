"""Drop-in import surface: the reference's models do
`from libs.GANet.modules.GANet import ...` and
`from libs.sync_bn.modules.sync_bn import BatchNorm2d, BatchNorm3d`
(models/GANet_deep.py:4-8).  Putting this repository before the reference on
sys.path makes those imports resolve to the B200-native implementation."""
