"""`from libs.GANet.build.lib import GANet` (reference functions/GANet.py:3) resolves to the
reference's native module surface implemented over the C ABI: the compiled pybind11 module
`ganet_b200/lib/GANet*.so` (csrc/ganet_pybind.cpp) when it has been built
(`python -m ganet_b200.build --pybind`), else the same six entry points in Python
(ganet_b200.legacy_native).  Both drive libganet_b200.so; neither computes anything itself."""
import glob
import importlib.util
import os

_so = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "..",
                                    "ganet_b200", "lib", "GANet*.so")))
if _so:
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    _spec = importlib.util.spec_from_file_location("GANet", _so[0])
    GANet = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(GANet)
else:
    from ganet_b200 import legacy_native as GANet  # noqa: F401
