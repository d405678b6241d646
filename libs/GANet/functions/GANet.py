"""Reference import path libs/GANet/functions/GANet.py -> ganet_b200.functions."""
from ganet_b200.functions import (  # noqa: F401
    SgaFunction, LgaFunction, Lga2Function, Lga3Function, Lga3dFunction, Lga3d2Function,
    Lga3d3Function, Lgf2Function, MyLossFunction, MyLoss2Function)
from ganet_b200 import legacy_native as GANet  # noqa: F401  (reference: `from ..build.lib import GANet`)
