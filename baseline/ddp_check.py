"""BASELINE.json config 4, correctness side: one GANet training step under DistributedDataParallel
(one process per GPU, NCCL gradient all-reduce, nn.SyncBatchNorm) against the same step on ONE GPU
with the whole batch ("Distributed": per-rank loss equals the single-GPU
per-sample loss; the all-reduced gradients equal those of the full batch).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 baseline/ddp_check.py --model GANet_deep --height 240 --width 624

The model is the reference's own (baseline/refmodels.py) on the new operators; loss as train.py:116-118
(SceneFlow branch).  Rank 0 prints one JSON line and exits non-zero if a tolerance is missed.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def loss_fn(F, outs, target):
    if len(outs) == 3:                      # GANet_deep: train.py:118
        d0, d1, d2 = outs
        return 0.2 * F.smooth_l1_loss(d0, target) + 0.6 * F.smooth_l1_loss(d1, target) + F.smooth_l1_loss(d2, target)
    d1, d2 = outs                           # GANet11: train.py:112
    return 0.4 * F.smooth_l1_loss(d1, target) + 1.2 * F.smooth_l1_loss(d2, target)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="GANet_deep")
    ap.add_argument("--height", type=int, default=240)
    ap.add_argument("--width", type=int, default=624)
    ap.add_argument("--rtol-loss", type=float, default=1e-4)
    ap.add_argument("--rtol-grad", type=float, default=2e-3)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from torch.nn.parallel import DistributedDataParallel as DDP
    from baseline import refmodels

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = False

    gen = torch.Generator(device="cpu").manual_seed(11)
    left = torch.randn(world, 3, a.height, a.width, generator=gen)
    right = torch.randn(world, 3, a.height, a.width, generator=gen)
    target = torch.rand(world, a.height, a.width, generator=gen) * 191.0

    # ---- DDP: rank r owns sample r ---------------------------------------------------
    model = refmodels.build(a.model, 192, seed=5, device=dev)
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    ddp = DDP(torch.nn.SyncBatchNorm.convert_sync_batchnorm(model), device_ids=[local])
    ddp.train()
    outs = ddp(left[rank:rank + 1].to(dev), right[rank:rank + 1].to(dev))
    loss = loss_fn(F, outs, target[rank:rank + 1].to(dev))
    loss.backward()
    losses = [torch.zeros((), device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    grads_ddp = {n: p.grad.detach().clone() for n, p in ddp.module.named_parameters() if p.grad is not None}
    # every rank must hold the same all-reduced gradient
    chk = torch.stack([g.double().sum() for g in grads_ddp.values()]).sum()
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same_on_all_ranks = bool(lo == hi)

    # ---- one GPU, whole batch, plain BatchNorm (statistics over the batch) ------------
    ok, line = True, None
    if rank == 0:
        single = refmodels.build(a.model, 192, seed=5, device=dev)
        single.load_state_dict(state0)
        single.train()
        outs1 = single(left.to(dev), right.to(dev))
        # per-sample losses of the single-GPU run, and the batch loss whose gradient DDP's mean equals
        per_sample = [loss_fn(F, [o[i:i + 1] for o in outs1], target[i:i + 1].to(dev)) for i in range(world)]
        total = sum(per_sample) / world
        total.backward()
        worst_loss = max(abs(float(losses[i]) - float(per_sample[i])) / abs(float(per_sample[i]))
                         for i in range(world))
        worst_grad, worst_name = 0.0, None
        named = dict(single.named_parameters())
        for n, g in grads_ddp.items():
            ref = named[n.replace("module.", "")].grad
            scale = max(float(ref.abs().max()), 1e-30)
            err = float((g - ref).abs().max()) / scale
            if err > worst_grad:
                worst_grad, worst_name = err, n
        ok = same_on_all_ranks and worst_loss <= a.rtol_loss and worst_grad <= a.rtol_grad
        line = {"check": "ddp_step", "model": a.model, "world": world, "height": a.height, "width": a.width,
                "per_rank_loss": [float(v) for v in losses],
                "single_gpu_per_sample_loss": [float(v) for v in per_sample],
                "worst_loss_rel_err": worst_loss, "worst_grad_rel_err": worst_grad,
                "worst_grad_param": worst_name, "n_param_tensors": len(grads_ddp),
                "grads_identical_on_all_ranks": same_on_all_ranks,
                "sync_bn_layers": sum(isinstance(m, torch.nn.SyncBatchNorm) for m in ddp.modules()),
                "rtol_loss": a.rtol_loss, "rtol_grad": a.rtol_grad, "ok": ok}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
