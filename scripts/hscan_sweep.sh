#!/bin/bash
# development aid: headline-shape timings of the horizontal kernels under different stage-ring depths
for s in 2 3 4; do
  echo "== GANET_HSCAN_STAGES=$s"; GANET_HSCAN_STAGES=$s timeout 200 python scripts/microbench.py --iters 3 --directions-only | grep -E "right|left"
done
for s in 1 2; do
  echo "== GANET_HSCAN_BWD_STAGES=$s"; GANET_HSCAN_BWD_STAGES=$s timeout 200 python scripts/microbench.py --iters 3 | grep -E "kept|backward"
done
