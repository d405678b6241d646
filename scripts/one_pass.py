#!/usr/bin/env python
"""One SGA fwd+bwd and one LGA2 fwd+bwd on a single sample -- the command profiled by ncu."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ganet_b200 import ops
shape = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else [1, 32, 192, 240, 624]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
N, C, D, H, W = shape
dev = torch.device("cuda")
x = torch.randn(N, C, D, H, W, device=dev); go = torch.randn_like(x)
g = [F.normalize(torch.randn(N, C, 5, H, W, device=dev), p=1, dim=2) for _ in range(4)]
xl = torch.randn(N, D, H, W, device=dev); gol = torch.randn_like(xl)
fl = F.normalize(torch.randn(N, 75, H, W, device=dev), p=1, dim=1)
keep = ops.keep_aggregates_policy(x, True) and os.environ.get("ONE_PASS_NO_KEEP") is None
print("aggregates kept for backward:", keep)
for _ in range(reps):
    if keep:
        out, mask, agg = ops.sga_forward(x, *g, keep_aggregates=True)
        gi, gg = ops.sga_backward(x, *g, mask, go, aggregates=agg)
    else:
        out, mask = ops.sga_forward(x, *g)
        gi, gg = ops.sga_backward(x, *g, mask, go)
    y1 = ops.lga_forward(xl, fl, 2); y = ops.lga_forward(y1, fl, 2)
    g1, gf = ops.lga_backward(y1, fl, gol, 2); gx, gf = ops.lga_backward(xl, fl, g1, 2, grad_f=gf)
torch.cuda.synchronize()
print("done")
