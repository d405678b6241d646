"""Loader for oracle/_ref/GANet*.so -- the UNMODIFIED reference CUDA extension
(libs/GANet/setup.py, built for sm_100a by oracle/build_ref.py).  Needs a GPU to
run.  TEST INFRASTRUCTURE ONLY.

The helpers restate the reference's Function-level call sequences
(libs/GANet/functions/GANet.py) with its buffer contract: the caller zero-fills
every output and scratch tensor.
"""
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def so_path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "GANet*.so")))
    return hits[0] if hits else None


def available():
    return so_path() is not None


def module():
    global _mod
    if _mod is None:
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        path = so_path()
        if path is None:
            raise RuntimeError("oracle/_ref/GANet*.so missing: run `python oracle/build_ref.py` "
                               "where /root/reference exists")
        spec = importlib.util.spec_from_file_location("GANet", path)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def sga_forward(x, g0, g1, g2, g3):
    """SgaFunction.forward (functions/GANet.py:10-22) -> out, mask (fp32), temp_out"""
    import torch
    out = torch.zeros_like(x); temp = torch.zeros_like(x); mask = torch.zeros_like(x)
    module().sga_cuda_forward(x, g0, g1, g2, g3, temp, out, mask)
    return out, mask, temp


def sga_backward(x, g0, g1, g2, g3, temp_out, mask, grad_out):
    """SgaFunction.backward (functions/GANet.py:24-48) -> grad_in, grads, max_idx (fp32)"""
    import torch
    N, C, D, H, W = x.shape
    gi = torch.zeros_like(x); tg = torch.zeros_like(x)
    gg = [torch.zeros_like(g0) for _ in range(4)]
    idx = torch.zeros((N, C, H, W), dtype=x.dtype, device=x.device)
    module().sga_cuda_backward(x, g0, g1, g2, g3, temp_out.clone(), mask, idx, grad_out, tg, gi,
                               gg[0], gg[1], gg[2], gg[3])
    return gi, tuple(gg), idx


def _fns(x):
    m = module()
    return (m.lga_cuda_forward, m.lga_cuda_backward) if x.dim() == 4 else \
        (m.lga3d_cuda_forward, m.lga3d_cuda_backward)


def lga_forward(x, f, radius=2):
    import torch
    y = torch.zeros_like(x)
    _fns(x)[0](x, f, y, radius)
    return y


def lga_backward(x, f, grad_out, grad_f=None, radius=2):
    import torch
    gx = torch.zeros_like(x)
    gf = torch.zeros_like(f) if grad_f is None else grad_f
    _fns(x)[1](x, f, grad_out, gx, gf, radius)
    return gx, gf


def lga2_forward(x, f, radius=2):
    """Lga2Function.forward (functions/GANet.py:176-187) -> y, y1"""
    y1 = lga_forward(x, f, radius)
    return lga_forward(y1, f, radius), y1


def lga2_backward(x, f, y1, grad_out, radius=2):
    """Lga2Function.backward (functions/GANet.py:189-203) -> gx, gf"""
    g1, gf = lga_backward(y1, f, grad_out, None, radius)
    gx, gf = lga_backward(x, f, g1, gf, radius)
    return gx, gf
