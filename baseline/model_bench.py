"""BASELINE.json configs 2-4: the reference's own models (baseline/refmodels.py: models/GANet11.py,
models/GANet_deep.py, unmodified) running on the new operators.

    python bench.py --config 2            GANet-11   inference, 240x624,  max_disp 192, one image per GPU
    python bench.py --config 3            GANet-deep inference, 384x1248, max_disp 192, one image per GPU
    torchrun ... bench.py --config 4 --gpus 8
                                          GANet-deep training step (train.py:114-123: forward, SceneFlow
                                          loss, backward, Adam step) under DistributedDataParallel +
                                          nn.SyncBatchNorm, one 240x624 sample per GPU, NCCL all-reduce of
                                          the 26.3 MB of gradients

Rank 0 prints ONE JSON line: ms per image (per training step), device-timed with CUDA events, max over
ranks; `e2e` = the same loop starting from pinned host images with the disparity map (the loss)
read back to the host every step; `hot_path` = the share of the step spent in this repository's
kernels, `nccl_ms` = the gradient all-reduce kernels, both from one profiled extra step;
`reference_cuda_on_this_gpu` = the same model object on the UNMODIFIED reference extension
(baseline/refops.py) on the same GPU.  Synthetic inputs: randn image pairs (the loaders produce
per-channel zero-mean / unit-std images, dataloader/dataset.py:136-144), target = rand*191, weights from
the models' own initialisation under a fixed seed.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    2: dict(model="GANet11", H=240, W=624, train=False,
            metric="GANet-11 inference latency per image, 240x624, max_disp=192"),
    3: dict(model="GANet_deep", H=384, W=1248, train=False,
            metric="GANet-deep inference latency per image, 384x1248 (KITTI full-res), max_disp=192"),
    4: dict(model="GANet_deep", H=240, W=624, train=True,
            metric="GANet-deep training step (fwd + loss + bwd + Adam), 240x624, max_disp=192, DDP batch-sharded"),
}

OUR_KERNEL_MARKS = ("sga_", "lga_", "merge4", "transpose_", "cost_volume", "disp_regression", "max_depth",
                    "ganet::")


def _loss(F, outs, target):
    from baseline.ddp_check import loss_fn
    return loss_fn(F, outs, target)


def _profile_one_step(torch, step_fn):
    """Kernel-time breakdown of one step (CUPTI through torch.profiler): ours / nccl / everything else."""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step_fn()
            torch.cuda.synchronize()
        ours = nccl = other = 0.0
        n_ours = 0
        per = {}
        for ev in prof.events():
            if getattr(ev, "device_type", None) is None or "cuda" not in str(ev.device_type).lower():
                continue
            us = float(getattr(ev, "device_time", 0.0) or getattr(ev, "cuda_time", 0.0) or 0.0)
            if us <= 0.0:
                continue
            name = ev.name
            if "nccl" in name.lower():
                nccl += us
            elif any(m in name for m in OUR_KERNEL_MARKS) and "cudnn" not in name and "at::" not in name:
                ours += us
                n_ours += 1
                key = name.split("<")[0].split("(")[0]
                per[key] = per.get(key, 0.0) + us
            else:
                other += us
        tot = ours + nccl + other
        if tot <= 0:
            return None
        top = sorted(per.items(), key=lambda kv: -kv[1])[:8]
        return {"ours_ms": ours / 1e3, "nccl_ms": nccl / 1e3, "other_ms": other / 1e3,
                "ours_share_of_kernel_time": ours / tot, "our_launches": n_ours,
                "our_kernels_ms": {k: v / 1e3 for k, v in top}}
    except Exception as exc:      # noqa: BLE001  (profiling is evidence, not the measurement)
        return {"unavailable": "%s: %s" % (type(exc).__name__, exc)}


def run(a):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from baseline import refmodels
    sys.path.insert(0, ROOT)
    from bench import ClockSampler, dist_setup, max_over_ranks

    cfg = CONFIGS[a.config]
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --config: no CUDA device; the product path has no CPU fallback")
    if not refmodels.available():
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"config_id": a.config, "unavailable": "baseline/_ref/models missing (run "
                              "python oracle/build_ref.py where /root/reference exists)"}))
        return 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    world, rank, local = dist_setup("nccl", dev)
    torch.backends.cudnn.benchmark = True
    H, W, train = cfg["H"], cfg["W"], cfg["train"]
    nb = max(1, a.per_gpu_batch)

    model = refmodels.build(cfg["model"], 192, seed=0, device=dev)
    n_params = sum(p.numel() for p in model.parameters())
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    h_left = torch.randn(nb, 3, H, W, generator=gen).pin_memory()
    h_right = torch.randn(nb, 3, H, W, generator=gen).pin_memory()
    h_target = (torch.rand(nb, H, W, generator=gen) * 191.0).pin_memory()
    d_left, d_right, d_target = h_left.to(dev), h_right.to(dev), h_target.to(dev)
    h_out = torch.empty(nb, H, W).pin_memory()
    h_loss = torch.empty(()).pin_memory()

    if train:
        net = model
        if world > 1:
            net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
            # find_unused_parameters: the models define layers their forward never calls (deconv0b)
            net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local],
                                                            find_unused_parameters=True)
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.999))      # train.py:74

        def step(left, right, target):
            opt.zero_grad(set_to_none=True)
            loss = _loss(F, net(left, right), target)
            loss.backward()
            opt.step()
            return loss.detach()
    else:
        net = model
        # running statistics as a trained checkpoint would have them (see tests/test_gpu_models.py)
        for m in net.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 1.0
        net.train()
        with torch.no_grad():
            net(d_left, d_right)
        net.eval()

        def step(left, right, target):
            with torch.no_grad():
                return net(left, right)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync_all()
        return max_over_ranks(e0.elapsed_time(e1), world, dev) / n

    def resident():
        return step(d_left, d_right, d_target)

    def from_host():
        left = h_left.to(dev, non_blocking=True)
        right = h_right.to(dev, non_blocking=True)
        target = h_target.to(dev, non_blocking=True) if train else None
        r = step(left, right, target)
        (h_loss if train else h_out).copy_(r, non_blocking=True)

    for _ in range(max(3, a.warmup)):
        resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(resident, a.steps)
    clk = clocks.stop() if rank == 0 else None
    from_host()
    ms_e2e = timed(from_host, a.steps)
    prof = _profile_one_step(torch, resident)
    mem_gb = torch.cuda.max_memory_allocated(dev) / 1e9

    # standalone NCCL all-reduce of a gradient-sized buffer (what DDP moves per step)
    allreduce_ms = None
    if train and world > 1:
        flat = torch.zeros(n_params, device=dev)
        for _ in range(3):
            dist.all_reduce(flat)
        allreduce_ms = timed(lambda: dist.all_reduce(flat), 10)

    fused = None
    if getattr(a, "fuse_sga_blocks", False):
        from ganet_b200.fused import fuse_sga_blocks, unfuse_sga_blocks
        nblk = fuse_sga_blocks(model)
        resident()
        fms = timed(resident, a.steps)
        unfuse_sga_blocks(model)
        fused = {"ms": fms if train else fms / nb, "sga_blocks_fused": nblk,
                 "note": "SGABlock prologue (split + 4x F.normalize) as one kernel each way, SURVEY.md 8f-2"}

    ref = None
    if not a.no_ref_gpu and world == 1:
        try:
            from baseline import refops
            with refops.reference_ops(model):
                resident()
                ref_ms = timed(resident, max(1, min(3, a.steps)))
            ref = {"ms": ref_ms, "speedup": ref_ms / ms,
                   "note": "same model object and weights, hot-path modules on oracle/_ref/GANet*.so"}
        except Exception as exc:      # noqa: BLE001
            ref = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        unit = "ms/step" if train else "ms/image"
        per = ms if train else ms / nb
        per_e2e = ms_e2e if train else ms_e2e / nb
        h2d = (h_left.numel() + h_right.numel() + (h_target.numel() if train else 0)) * 4 * world
        d2h = (4 if train else h_out.numel() * 4) * world
        line = {
            "metric": cfg["metric"], "value": per, "unit": unit, "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE.json config %d: %s, %dx%d, max_disp 192, %d sample(s) per GPU"
                                   % (a.config, cfg["model"], H, W, nb),
                       "model_source": "reference models/%s.py unmodified (baseline/_ref/models), "
                                       "%d parameters" % (cfg["model"], n_params),
                       "global_batch": nb * world,
                       "parallelism": ("DDP x%d + SyncBatchNorm, NCCL gradient all-reduce" % world) if train
                       else "one image per GPU, no collective",
                       "l2": "activations of one step exceed L2 (cost volume 277-886 MB)"},
            "throughput": {"value": 1e3 * nb * world / ms, "unit": "steps/s" if train else "images/s"},
            "e2e": {"value": per_e2e, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "model(left, right) from pinned host images; %s read back to the host every step"
                           % ("loss" if train else "disparity map")},
            "hot_path": prof, "gpu_launches": (prof or {}).get("our_launches"), "clocks": clk,
            "peak_memory_gb": mem_gb, "fused_sga_blocks": fused, "reference_cuda_on_this_gpu": ref,
        }
        if train:
            line["nccl"] = {"grad_bytes": n_params * 4, "standalone_allreduce_ms": allreduce_ms,
                            "allreduce_kernels_ms_in_step": (prof or {}).get("nccl_ms")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0
