#!/usr/bin/env python
"""Per-call device timings of the C-ABI entry points (development aid).
    python scripts/microbench.py [--shape N C D H W] [--iters 5]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganet_b200 import ops  # noqa: E402


def timeit(fn, iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_lga(a, N, D, H, W, dev):
    xl = torch.randn(N, D, H, W, device=dev)
    fl = F.normalize(torch.randn(N, 75, H, W, device=dev), p=1, dim=1)
    gol = torch.randn_like(xl)
    ms = timeit(lambda: ops.lga_forward(xl, fl, 2), a.iters)
    print("  lga_forward (1 pass)  %8.3f ms  %7.1f Gvox/s" % (ms, xl.numel() / ms / 1e6))
    ms = timeit(lambda: ops.lga_backward(xl, fl, gol, 2), a.iters)
    print("  lga_backward (1 pass) %8.3f ms  %7.1f Gvox/s" % (ms, xl.numel() / ms / 1e6))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=5, default=[1, 32, 192, 240, 624])
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--lga", action="store_true")
    ap.add_argument("--directions-only", action="store_true", help="time the four raw scans and stop")
    ap.add_argument("--lga-only", action="store_true", help="time one LGA pass each way and stop")
    a = ap.parse_args()
    N, C, D, H, W = a.shape
    dev = torch.device("cuda")
    if a.lga_only:
        time_lga(a, N, D, H, W, dev)
        return
    x = torch.randn(N, C, D, H, W, device=dev)
    go = torch.randn_like(x)
    g = [F.normalize(torch.randn(N, C, 5, H, W, device=dev), p=1, dim=2) for _ in range(4)]
    V = x.numel()
    print("shape", a.shape, "voxels %.1fM" % (V / 1e6))
    for d, name in enumerate(["down", "up", "right", "left"]):
        ms = timeit(lambda: ops.sga_direction(x, g[d], d), a.iters)
        print("  direction %-5s raw  %8.3f ms  %7.1f Gvox/s  %7.1f GB/s (8 B/vox)" % (name, ms, V / ms / 1e6, 8 * V / ms / 1e6))
    if a.directions_only:
        return
    ms = timeit(lambda: ops.sga_forward(x, *g), a.iters)
    print("  sga_forward         %8.3f ms  %7.1f Gvox/s  algorithmic %.1f GB/s" % (ms, V / ms / 1e6, (9 + 80 / D) * V / ms / 1e6))
    out, mask = ops.sga_forward(x, *g)
    ms = timeit(lambda: ops.sga_backward(x, *g, mask, go), a.iters)
    print("  sga_backward        %8.3f ms  %7.1f Gvox/s  algorithmic %.1f GB/s" % (ms, V / ms / 1e6, (13 + 160 / D) * V / ms / 1e6))
    del out, mask
    msf = timeit(lambda: ops.sga_forward(x, *g, keep_aggregates=True), a.iters)
    out, mask, agg = ops.sga_forward(x, *g, keep_aggregates=True)
    msb = timeit(lambda: ops.sga_backward(x, *g, mask, go, aggregates=agg), a.iters)
    print("  kept aggregates: fwd %8.3f ms  bwd %8.3f ms  sum %8.3f ms" % (msf, msb, msf + msb))
    if a.lga:
        time_lga(a, N, D, H, W, dev)


if __name__ == "__main__":
    main()
