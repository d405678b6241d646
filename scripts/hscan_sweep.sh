#!/bin/bash
# development aid: headline-shape timings under different L2-prefetch distances / ring depths
run() { echo "== $*"; env "$@" timeout 300 python scripts/microbench.py --iters 3 | grep -E "direction|kept|backward|forward"; }
run GANET_HSCAN_PREFETCH=0 GANET_HSCAN_BWD_PREFETCH=0
run GANET_HSCAN_PREFETCH=4 GANET_HSCAN_BWD_PREFETCH=4
run GANET_HSCAN_PREFETCH=8 GANET_HSCAN_BWD_PREFETCH=8
run GANET_HSCAN_PREFETCH=8 GANET_HSCAN_BWD_PREFETCH=16 GANET_HSCAN_STAGES=2
run GANET_VERT_PREFETCH=2
run GANET_VERT_PREFETCH=4
run GANET_VERT_PREFETCH=8
