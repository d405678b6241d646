#!/usr/bin/env python
"""BASELINE.json config 5: SGA fwd+bwd and LGA2 fwd+bwd sweep D in {96,192,288} x HxW in {240x624,
480x1248}, one sample (C=32) on one GPU; prints voxels/s, the fraction of the HBM roofline for SGA
(22 + 240/D B/voxel) and the combined SGA+LGA rate as bench.py defines it."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganet_b200 import ops  # noqa: E402

peak = 6489.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for (H, W) in ((240, 624), (480, 1248)):
    for D in (96, 192, 288):
        C = 32
        x = torch.randn(1, C, D, H, W, device="cuda")
        go = torch.randn_like(x)
        g = [F.normalize(torch.randn(1, C, 5, H, W, device="cuda"), p=1, dim=2) for _ in range(4)]
        keep = ops.keep_aggregates_policy(x, True)

        def once():
            if keep:
                out, mask, agg = ops.sga_forward(x, *g, keep_aggregates=True)
                ops.sga_backward(x, *g, mask, go, aggregates=agg)
            else:
                out, mask = ops.sga_forward(x, *g)
                ops.sga_backward(x, *g, mask, go)
        once(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            once()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        V = x.numel()
        gbs = (22 + 240.0 / D) * V / ms / 1e6
        del x, go, g
        torch.cuda.empty_cache()
        xl = torch.randn(1, D, H, W, device="cuda")
        fl = F.normalize(torch.randn(1, 75, H, W, device="cuda"), p=1, dim=1)
        gol = torch.randn_like(xl)

        def lga_once():
            y1 = ops.lga_forward(xl, fl, 2); ops.lga_forward(y1, fl, 2)
            g1, gf = ops.lga_backward(y1, fl, gol, 2); ops.lga_backward(xl, fl, g1, 2, gf)
        lga_once(); torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            lga_once()
        e1.record(); torch.cuda.synchronize()
        lga_ms = e0.elapsed_time(e1) / 5
        rows.append({"D": D, "H": H, "W": W, "voxels": V, "ms": ms, "gvox_s": V / ms / 1e6,
                     "alg_gbs": gbs, "frac": gbs / peak, "kept_aggregates": bool(keep),
                     "lga2_voxels": xl.numel(), "lga2_ms": lga_ms, "lga2_gvox_s": xl.numel() / lga_ms / 1e6,
                     "combined_gvox_s": (V + xl.numel()) / (ms + lga_ms) / 1e6})
        x = go = g = None
        del xl, fl, gol
        print(rows[-1]); sys.stdout.flush()
        del x, go, g
        torch.cuda.empty_cache()
json.dump(rows, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "sweep.json"), "w"), indent=1)
