"""Reference import path libs/GANet/modules/GANet.py -> ganet_b200.modules."""
from ganet_b200.modules import (  # noqa: F401
    SGA, LGA, LGA2, LGA3, LGA3D, LGA3D2, LGA3D3, DisparityRegression, GetCostVolume, MyLoss,
    MyLoss2, MyNormalize)
from ganet_b200.functions import (  # noqa: F401
    LgaFunction, Lga2Function, Lga3Function, Lga3dFunction, Lga3d2Function, Lga3d3Function,
    MyLossFunction, MyLoss2Function, SgaFunction)
