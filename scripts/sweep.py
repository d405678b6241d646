#!/usr/bin/env python
"""BASELINE.json config 5: SGA fwd+bwd and LGA2 fwd+bwd sweep D in {96,192,288} x HxW in {240x624,
480x1248}, batch 8 (C=32) sharded over the ranks (torchrun; one process per GPU, no data-path
collective), as bench.py shards it.  Per cell: whole-job voxels/s (all ranks, max-over-ranks device
time), the fraction of the HBM roofline for SGA (22 + 240/D B/voxel) and the combined SGA+LGA rate as
bench.py defines it.

    python scripts/sweep.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29540 scripts/sweep.py                  # eight

A rank walks its share of the 8 samples one at a time through ONE set of device buffers (a sample is
1.8-22 GB per tensor, far larger than L2, so reusing the buffers does not warm anything)."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import dist_setup, max_over_ranks, shard_samples  # noqa: E402
from ganet_b200 import ops  # noqa: E402

B, C = 8, 32
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
world, rank, local = dist_setup("nccl", dev)
peak = 6489.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
cells = [(H, W, D) for (H, W) in ((240, 624), (480, 1248)) for D in (96, 192, 288)]
if len(sys.argv) > 1:                       # e.g. "240x624x192,480x1248x96"
    cells = [tuple(int(v) for v in c.split("x")) for c in sys.argv[1].split(",")]
rows = []
ev = lambda: torch.cuda.Event(enable_timing=True)     # noqa: E731
for (H, W, D) in cells:
    n_mine = len(shard_samples(B, world, rank))
    x = torch.randn(1, C, D, H, W, device=dev)
    go = torch.randn_like(x)
    g = [F.normalize(torch.randn(1, C, 5, H, W, device=dev), p=1, dim=2) for _ in range(4)]
    keep = ops.keep_aggregates_policy(x, True)

    def once():
        if keep:
            out, mask, agg = ops.sga_forward(x, *g, keep_aggregates=True)
            ops.sga_backward(x, *g, mask, go, aggregates=agg)
        else:
            out, mask = ops.sga_forward(x, *g)
            ops.sga_backward(x, *g, mask, go)

    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n_mine):
        once()
    e1.record(); torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    V = B * x.numel()
    gbs = (22 + 240.0 / D) * V / ms / 1e6
    del x, go, g
    torch.cuda.empty_cache()
    xl = torch.randn(1, D, H, W, device=dev)
    fl = F.normalize(torch.randn(1, 75, H, W, device=dev), p=1, dim=1)
    gol = torch.randn_like(xl)

    def lga_once():
        y1 = ops.lga_forward(xl, fl, 2); ops.lga_forward(y1, fl, 2)
        g1, gf = ops.lga_backward(y1, fl, gol, 2); ops.lga_backward(xl, fl, g1, 2, gf)

    lga_once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(n_mine):
        lga_once()
    e1.record(); torch.cuda.synchronize()
    lga_ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    VL = B * xl.numel()
    rows.append({"n_gpus": world, "D": D, "H": H, "W": W, "batch": B, "voxels": V, "sga_ms": ms,
                 "sga_gvox_s": V / ms / 1e6, "sga_alg_gbs_all_gpus": gbs, "sga_frac_of_hbm_roofline": gbs / (peak * world),
                 "kept_aggregates": bool(keep),
                 "lga2_voxels": VL, "lga2_ms": lga_ms, "lga2_gvox_s": VL / lga_ms / 1e6,
                 "combined_gvox_s": (V + VL) / (ms + lga_ms) / 1e6})
    del xl, fl, gol
    torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(rows[-1])); sys.stdout.flush()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_n%d.json" % world), "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
