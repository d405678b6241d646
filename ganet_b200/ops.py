"""Tensor-level wrappers over the C ABI (no autograd).  Every function takes and
returns contiguous fp32 CUDA tensors, allocates its outputs with torch (device
memory plumbing only) and enqueues on the current CUDA stream."""
import os

import torch

from . import _lib
from ._lib import check, ptr, stream

_i64 = _lib._i64
_int = _lib._int


def _dims5(x):
    if x.dim() != 5:
        raise ValueError("expected a 5-D (N,C,D,H,W) tensor, got %s" % (tuple(x.shape),))
    return [_i64(int(v)) for v in x.shape]


def _workspace_budget(device=None):
    """Scratch budget per call: GANET_B200_WORKSPACE_BYTES, else a quarter of the device
    memory (45 GB on a 180 GB B200) -- enough for a 920M-voxel sample in one chunk."""
    env = os.environ.get("GANET_B200_WORKSPACE_BYTES")
    if env:
        return int(env)
    return torch.cuda.get_device_properties(device).total_memory // 4


def _free_bytes(device):
    """Bytes a new allocation can count on: free device memory plus what torch's caching
    allocator holds but has not handed out."""
    free, _ = torch.cuda.mem_get_info(device)
    cached = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return int(free + max(0, cached))


def _workspace(x, ws_min, ws_best, workspace_bytes):
    """Scratch for one call: as much of `best` as the budget and the free memory allow, never below
    `min` (the native side walks the (n,c) slices in chunks that fit).  If even that allocation
    fails, retry with the minimum before giving up."""
    if workspace_bytes is None:
        budget = _workspace_budget(x.device)
        if ws_best > (1 << 30):          # cudaMemGetInfo costs a driver round trip: only for large requests
            budget = min(budget, _free_bytes(x.device) // 2)
    else:
        budget = int(workspace_bytes)
    nbytes = int(min(ws_best, max(ws_min, budget)))
    try:
        return torch.empty(nbytes, dtype=torch.uint8, device=x.device), nbytes
    except torch.cuda.OutOfMemoryError:
        if nbytes <= ws_min:
            raise
        torch.cuda.empty_cache()
        return torch.empty(int(ws_min), dtype=torch.uint8, device=x.device), int(ws_min)


def aggregate_volumes(x):
    """How many x-sized fp32 volumes sga_forward(keep_aggregates=True) keeps for this shape (4 when
    the horizontal scans run in the standard layout, 5 on the transposed path)."""
    return int(_lib.lib().ganet_sga_aggregate_volumes(*_dims5(x)))


def keep_aggregates_policy(x, needs_backward=True):
    """Should forward keep the directional aggregates (16-20 bytes per voxel) so that backward can skip
    its recompute passes?  GANET_B200_KEEP_AGGREGATES = 0 | 1 | auto (default auto: yes when a
    backward will follow, D <= 288 and twice the buffer is available to the allocator -- free
    device memory plus torch's cached-but-unused blocks).  GANET_B200_KEEP_AGGREGATES_BUDGET caps, in
    bytes, what all live SgaFunction nodes of the process may hold this way (default: no cap)."""
    mode = os.environ.get("GANET_B200_KEEP_AGGREGATES", "auto")
    if mode == "0" or not needs_backward or x.shape[2] > 288:
        return False
    if mode == "1":
        return True
    need = 4 * aggregate_volumes(x) * x.numel()
    cap = os.environ.get("GANET_B200_KEEP_AGGREGATES_BUDGET")
    if cap is not None and _kept_bytes[0] + need > int(cap):
        return False
    if need <= (1 << 30):
        # small calls (the models' shapes: 0.1-1.8 GB of aggregates... up to 1 GB here) skip the driver
        # query -- it costs more than the call; an allocation failure falls back to recomputing
        return True
    return _free_bytes(x.device) >= 2 * need


_kept_bytes = [0]       # bytes of aggregates currently held by autograd nodes (functions.SgaFunction)


def track_kept_aggregates(agg):
    """Count a kept-aggregates buffer against GANET_B200_KEEP_AGGREGATES_BUDGET for as long as its storage lives."""
    import weakref
    nbytes = agg.numel() * agg.element_size()
    _kept_bytes[0] += nbytes

    def _release(n=nbytes):
        _kept_bytes[0] = max(0, _kept_bytes[0] - n)
    weakref.finalize(agg.untyped_storage(), _release)


def sga_forward(x, g0, g1, g2, g3, workspace_bytes=None, keep_aggregates=False):
    """-> out (N,C,D,H,W) f32, mask (N,C,D,H,W) u8 [, aggregates (5, N*C*D*H*W) f32]"""
    N, C, D, H, W = x.shape
    for g in (g0, g1, g2, g3):
        if tuple(g.shape) != (N, C, 5, H, W):
            raise ValueError("guidance must be (N,C,5,H,W), got %s" % (tuple(g.shape),))
    for t in (x, g0, g1, g2, g3):
        ptr(t)                      # device / dtype / layout checks before any allocation
    with torch.cuda.device_of(x):
        out = torch.empty_like(x)
        mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        L = _lib.lib()
        dims = _dims5(x)
        ws, ws_bytes = _workspace(x, L.ganet_sga_forward_workspace_min(*dims),
                                  L.ganet_sga_forward_workspace_best(*dims), workspace_bytes)
        agg = (torch.empty((L.ganet_sga_aggregate_volumes(*dims), x.numel()), dtype=x.dtype, device=x.device)
               if keep_aggregates else None)
        check(L.ganet_sga_forward(ptr(x), ptr(g0), ptr(g1), ptr(g2), ptr(g3), ptr(out),
                                  ptr(mask, torch.uint8), ptr(agg) if agg is not None else None,
                                  ptr(ws, torch.uint8), _lib._sz(ws_bytes), *dims, stream()))
    if keep_aggregates:
        return out, mask, agg
    return out, mask


def sga_direction(x, g, direction):
    """one directional aggregate (0 down, 1 up, 2 right, 3 left), no max-combine"""
    with torch.cuda.device_of(x):
        a = torch.empty_like(x)
        check(_lib.lib().ganet_sga_direction(ptr(x), ptr(g), ptr(a), _int(direction), *_dims5(x),
                                             stream()))
    return a


def sga_backward(x, g0, g1, g2, g3, mask, grad_out, want_max_idx=False, workspace_bytes=None,
                 aggregates=None):
    """-> grad_in, (gg0, gg1, gg2, gg3)[, max_idx int32 (N,C,H,W)]
    `aggregates`: the tensor sga_forward(..., keep_aggregates=True) returned, or None."""
    N, C, D, H, W = x.shape
    L = _lib.lib()
    for t in (x, g0, g1, g2, g3, grad_out):
        ptr(t)
    with torch.cuda.device_of(x):
        gi = torch.empty_like(x)
        gg = [torch.empty_like(g0) for _ in range(4)]
        idx = torch.empty((N, C, H, W), dtype=torch.int32, device=x.device) if want_max_idx else None
        dims = _dims5(x)
        ws, ws_bytes = _workspace(x, L.ganet_sga_backward_workspace_min(*dims),
                                  L.ganet_sga_backward_workspace_best(*dims), workspace_bytes)
        if aggregates is not None and tuple(aggregates.shape) != (L.ganet_sga_aggregate_volumes(*dims), x.numel()):
            raise ValueError("aggregates must be the tensor returned by sga_forward(keep_aggregates=True)")
        check(L.ganet_sga_backward(ptr(x), ptr(g0), ptr(g1), ptr(g2), ptr(g3),
                                   ptr(mask, torch.uint8),
                                   ptr(aggregates) if aggregates is not None else None,
                                   ptr(grad_out), ptr(gi), ptr(gg[0]),
                                   ptr(gg[1]), ptr(gg[2]), ptr(gg[3]),
                                   ptr(idx, torch.int32) if idx is not None else None,
                                   ptr(ws, torch.uint8), _lib._sz(ws_bytes), *dims, stream()))
    if want_max_idx:
        return gi, tuple(gg), idx
    return gi, tuple(gg)


def sga_guidance_forward(raw, channels):
    """SGABlock prologue: raw (N, 4*C*5, H, W) guidance -> four L1-normalised (N, C, 5, H, W) weight tensors
    (down, up, right, left), what models/GANet_deep.py:264-268 computes with split + view + F.normalize."""
    N, K, H, W = raw.shape
    C = int(channels)
    if K != 4 * C * 5:
        raise ValueError("guidance must have 4*C*5 = %d channels, got %d" % (4 * C * 5, K))
    with torch.cuda.device_of(raw):
        g = [torch.empty((N, C, 5, H, W), dtype=raw.dtype, device=raw.device) for _ in range(4)]
        check(_lib.lib().ganet_sga_guidance_forward(ptr(raw), ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]),
                                                    _i64(N), _i64(C), _i64(H), _i64(W), stream()))
    return tuple(g)


def sga_guidance_backward(raw, gg0, gg1, gg2, gg3):
    """SGABlock epilogue of backward: gradients w.r.t. the four normalised weight tensors -> gradient of raw."""
    N, K, H, W = raw.shape
    C = K // 20
    with torch.cuda.device_of(raw):
        out = torch.empty_like(raw)
        check(_lib.lib().ganet_sga_guidance_backward(ptr(raw), ptr(gg0), ptr(gg1), ptr(gg2), ptr(gg3), ptr(out),
                                                     _i64(N), _i64(C), _i64(H), _i64(W), stream()))
    return out


def _lga_dims(x, f, radius):
    if x.dim() not in (4, 5):
        raise ValueError("LGA input must be (N,D,H,W) or (N,C,D,H,W)")
    lead = 1
    for v in x.shape[:-3]:
        lead *= int(v)
    D, H, W = (int(v) for v in x.shape[-3:])
    F = 3 * (2 * radius + 1) ** 2
    if tuple(f.shape) != tuple(x.shape[:-3]) + (F, H, W):
        raise ValueError("LGA filters must be %s, got %s" % (tuple(x.shape[:-3]) + (F, H, W),
                                                            tuple(f.shape)))
    return _i64(lead), _i64(D), _i64(H), _i64(W)


def lga_forward(x, f, radius):
    with torch.cuda.device_of(x):
        y = torch.empty_like(x)
        check(_lib.lib().ganet_lga_forward(ptr(x), ptr(f), ptr(y), *_lga_dims(x, f, radius),
                                           _int(radius), stream()))
    return y


def lga_backward(x, f, grad_out, radius, grad_f=None):
    """-> grad_x, grad_f.  If grad_f is given it is accumulated into (+=)."""
    with torch.cuda.device_of(x):
        gx = torch.empty_like(x)
        acc = grad_f is not None
        gf = grad_f if acc else torch.empty_like(f)
        check(_lib.lib().ganet_lga_backward(ptr(x), ptr(f), ptr(grad_out), ptr(gx), ptr(gf),
                                            _int(int(acc)), *_lga_dims(x, f, radius), _int(radius),
                                            stream()))
    return gx, gf


def cost_volume_forward(x, y, maxdisp_plus1):
    N, C, H, W = x.shape
    if tuple(y.shape) != tuple(x.shape):
        raise ValueError("left/right feature maps must have the same shape")
    Dm = int(maxdisp_plus1)
    with torch.cuda.device_of(x):
        cost = torch.empty((N, 2 * C, Dm, H, W), dtype=x.dtype, device=x.device)
        check(_lib.lib().ganet_cost_volume_forward(ptr(x), ptr(y), ptr(cost), _i64(N), _i64(C),
                                                   _i64(Dm), _i64(H), _i64(W), stream()))
    return cost


def cost_volume_backward(grad_cost):
    N, C2, Dm, H, W = grad_cost.shape
    C = C2 // 2
    with torch.cuda.device_of(grad_cost):
        gx = torch.empty((N, C, H, W), dtype=grad_cost.dtype, device=grad_cost.device)
        gy = torch.empty_like(gx)
        check(_lib.lib().ganet_cost_volume_backward(ptr(grad_cost), ptr(gx), ptr(gy), _i64(N),
                                                    _i64(C), _i64(Dm), _i64(H), _i64(W), stream()))
    return gx, gy


def disp_regression_forward(p):
    N, Dm, H, W = p.shape
    with torch.cuda.device_of(p):
        out = torch.empty((N, H, W), dtype=p.dtype, device=p.device)
        check(_lib.lib().ganet_disp_regression_forward(ptr(p), ptr(out), _i64(N), _i64(Dm), _i64(H),
                                                       _i64(W), stream()))
    return out


def disp_regression_backward(grad_disp, Dm):
    N, H, W = grad_disp.shape
    with torch.cuda.device_of(grad_disp):
        gp = torch.empty((N, Dm, H, W), dtype=grad_disp.dtype, device=grad_disp.device)
        check(_lib.lib().ganet_disp_regression_backward(ptr(grad_disp), ptr(gp), _i64(N), _i64(Dm),
                                                        _i64(H), _i64(W), stream()))
    return gp


def norm_disp_regression_forward(x):
    """DispAgg tail: F.normalize(x, p=1, dim=1) + DisparityRegression in one pass -> disp (N,H,W), norm (N,H,W)."""
    N, Dm, H, W = x.shape
    with torch.cuda.device_of(x):
        disp = torch.empty((N, H, W), dtype=x.dtype, device=x.device)
        norm = torch.empty_like(disp)
        check(_lib.lib().ganet_norm_disp_regression_forward(ptr(x), ptr(disp), ptr(norm), _i64(N), _i64(Dm),
                                                            _i64(H), _i64(W), stream()))
    return disp, norm


def norm_disp_regression_backward(x, disp, norm, grad_disp):
    N, Dm, H, W = x.shape
    with torch.cuda.device_of(x):
        gx = torch.empty_like(x)
        check(_lib.lib().ganet_norm_disp_regression_backward(ptr(x), ptr(disp), ptr(norm), ptr(grad_disp), ptr(gx),
                                                             _i64(N), _i64(Dm), _i64(H), _i64(W), stream()))
    return gx
