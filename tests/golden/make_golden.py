#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own kernel bodies compiled for
the host (oracle/_ref/libganet_ref_cpu.so, built by oracle/build_ref.py from
/root/reference/libs/GANet/src/GANet_kernel.cu).  Needs /root/reference only at
build time; the committed vectors are what travels.

The reference ships no golden vectors, so these are the pin for
oracle/ganet_oracle.c (bit-exact with fused=0) and, with the fp32 tolerance the
parity tests state, for the CUDA kernels.

    python oracle/build_ref.py --cpu-only && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_cpu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

SGA_SHAPES = [(2, 3, 7, 5, 6), (1, 2, 1, 4, 3), (1, 1, 2, 1, 5), (1, 2, 3, 6, 1),
              (1, 2, 9, 8, 10), (1, 1, 33, 5, 7)]
LGA_SHAPES = [(2, 7, 5, 6), (1, 3, 9, 11), (1, 1, 1, 1), (1, 2, 3, 5, 4)]


def l1norm(a, axis):
    return (a / np.abs(a).sum(axis=axis, keepdims=True)).astype(np.float32)


def main():
    rng = np.random.default_rng(20260923)
    sga = {}
    for k, shape in enumerate(SGA_SHAPES):
        N, C, D, H, W = shape
        x = rng.standard_normal(shape).astype(np.float32)
        g = [l1norm(rng.standard_normal((N, C, 5, H, W)), 2) for _ in range(4)]
        go = rng.standard_normal(shape).astype(np.float32)
        out, mask, temp = ref_cpu.sga_forward(x, *g)
        gi, gg, idx = ref_cpu.sga_backward(x, *g, temp, mask, go)
        sga.update({f"x{k}": x, f"go{k}": go, f"out{k}": out, f"mask{k}": mask.astype(np.uint8),
                    f"left{k}": temp, f"gi{k}": gi, f"idx{k}": idx.astype(np.int32)})
        for d in range(4):
            sga[f"g{k}_{d}"] = g[d]
            sga[f"gg{k}_{d}"] = gg[d]
    np.savez_compressed(os.path.join(HERE, "sga_ref_cpu.npz"), n=len(SGA_SHAPES), **sga)

    lga = {}
    for k, shape in enumerate(LGA_SHAPES):
        x = rng.standard_normal(shape).astype(np.float32)
        fs = shape[:-3] + (75,) + shape[-2:]
        f = l1norm(rng.standard_normal(fs), len(fs) - 3)
        go = rng.standard_normal(shape).astype(np.float32)
        y, y1 = ref_cpu.lga2_forward(x, f, 2)
        gx, gf = ref_cpu.lga2_backward(x, f, y1, go, 2)
        lga.update({f"x{k}": x, f"f{k}": f, f"go{k}": go, f"y{k}": y, f"y1_{k}": y1,
                    f"gx{k}": gx, f"gf{k}": gf})
    np.savez_compressed(os.path.join(HERE, "lga_ref_cpu.npz"), n=len(LGA_SHAPES), **lga)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
