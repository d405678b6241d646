"""Loads the reference's model definitions (models/GANet_deep.py, models/GANet11.py), copied
byte for byte into baseline/_ref/models/ by oracle/build_ref.py, so that BASELINE.json's configs
2-4 run the reference's own consumers of the hot path.  Their imports
(`from libs.GANet.modules.GANet import ...`, `from libs.sync_bn.modules.sync_bn import ...`,
models/GANet_deep.py:4-8) resolve to this repository's drop-in `libs/` package, i.e. to the
sm_100a operators of ganet_b200 -- the models themselves are not touched."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = os.path.join(ROOT, "baseline", "_ref", "models")


def available():
    return os.path.exists(os.path.join(MODELS, "GANet_deep.py"))


def load(name):
    """name: 'GANet_deep' or 'GANet11' -> the module object (its class is `GANet`)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    path = os.path.join(MODELS, name + ".py")
    if not os.path.exists(path):
        raise RuntimeError("%s missing: run `python oracle/build_ref.py` where /root/reference exists" % path)
    modname = "ganet_reference_models." + name
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def build(name, maxdisp=192, seed=0, device=None):
    """The reference's `GANet(maxdisp)` with its own initialisation (models/GANet_deep.py:382-387)
    under a fixed seed."""
    import torch
    torch.manual_seed(seed)
    model = load(name).GANet(maxdisp)
    return model.to(device) if device is not None else model
