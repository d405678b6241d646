#!/bin/bash
# the 8-GPU evidence session (run under `gpurun --gpus 8`): headline bench, config-5 sweep, config 4, DDP check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2s_topo.txt 2>&1
timeout 600 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2s_bench_n4.json 2> gpurun_out/r2s_bench_n4.err
timeout 900 $TR --nproc-per-node 8 --master-port 29603 scripts/sweep.py > gpurun_out/r2s_sweep_n8.txt 2>&1
timeout 600 $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --config 4 --steps 10 --warmup 3 > gpurun_out/r2s_config4_n8.json 2> gpurun_out/r2s_config4_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29605 baseline/ddp_check.py --model GANet_deep --height 240 --width 624 2> gpurun_out/r2s_ddp_n8.err | grep "^{" > gpurun_out/r2s_ddp_check_n8.json
for f in gpurun_out/r2s_bench_n8.json gpurun_out/r2s_bench_n4.json gpurun_out/r2s_config4_n8.json gpurun_out/r2s_ddp_check_n8.json; do echo "== $f"; cut -c1-700 $f; done
cut -c1-260 gpurun_out/r2s_sweep_n8.txt
