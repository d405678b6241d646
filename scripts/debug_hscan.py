#!/usr/bin/env python
"""Development aid: isolate one SGA direction in backward by handing both implementations a mask that
selects only that direction, and report where the new kernels differ from the reference extension."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ganet_b200 import ops
from oracle import ref_gpu
from util import sga_inputs

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

shapes = [(1, 1, 12, 2, 16), (1, 1, 12, 3, 32), (1, 2, 24, 16, 48), (1, 1, 192, 4, 64), (2, 2, 65, 32, 80), (1, 8, 48, 48, 96)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1:6])]
for shape in shapes:
    x, g, go = sga_inputs(shape, seed=5)
    xt, gt, got = [torch.from_numpy(a).cuda() for a in (x,)][0], [torch.from_numpy(a).cuda() for a in g], torch.from_numpy(go).cuda()
    ro, rm, rtemp = ref_gpu.sga_forward(xt, *gt)
    for only in (None, 0, 1, 2, 3):
        if only is None:
            m8, mf = rm.to(torch.uint8), rm
        else:
            m8 = torch.full(shape, only, dtype=torch.uint8, device="cuda"); mf = m8.float()
        rgi, rgg, ridx = ref_gpu.sga_backward(xt, *gt, rtemp, mf, got)
        gi, gg, idx = ops.sga_backward(xt, *gt, m8, got, want_max_idx=True)
        torch.cuda.synchronize()
        errs = [rel(gi, rgi)] + [rel(gg[k], rgg[k]) for k in range(4)]
        print(shape, "only", only, "gi %.2e" % errs[0], "gg", " ".join("%.2e" % e for e in errs[1:]),
              "idx ok" if torch.equal(idx, ridx.to(torch.int32)) else "IDX DIFF")
        if errs[0] > 1e-4:
            d = (gi - rgi).abs()
            bad = (d > 1e-4 * rgi.abs().max()).nonzero()
            print("   gi bad count", bad.shape[0], "first", bad[:6].tolist())
            cols = torch.bincount(bad[:, 4], minlength=shape[4]).tolist()
            print("   bad per column", cols)
            deps = torch.bincount(bad[:, 2], minlength=shape[2]).tolist()
            print("   bad per depth", deps)
        for k in range(4):
            if errs[1 + k] > 1e-4:
                d = (gg[k] - rgg[k]).abs()
                bad = (d > 1e-4 * rgg[k].abs().max()).nonzero()
                print("   gg%d bad count" % k, bad.shape[0], "first", bad[:6].tolist(),
                      "per weight", torch.bincount(bad[:, 2], minlength=5).tolist())
