#!/usr/bin/env python
"""Training harness: the reference's train.py loop, one process per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
        harness/train_ddp.py --model GANet_deep --batchSize 8 --crop_height 240 --crop_width 624 --synthetic 64

Same flags as train.py:21-43 (batchSize is the GLOBAL batch, split over the ranks), same loss weights (:104-118),
Adam and learning-rate step (:74, :171-178), same checkpoint dictionary and file names (:164-169, :191-211);
state_dict keys carry the `module.` prefix the reference's DataParallel checkpoints have, and --resume accepts
checkpoints with or without it (strict=False, :75-82).  Differences: DistributedDataParallel + nn.SyncBatchNorm
over NCCL instead of nn.DataParallel + the vendored SyncBN; `--synthetic N` trains on N generated pairs
(harness/data.py) because the container has no dataset; the models are the reference's files
(baseline/_ref/models, or any directory given by --models_dir) on the operators of this repository."""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="GANet training on B200 (DDP)")
    p.add_argument("--crop_height", type=int, required=True)
    p.add_argument("--crop_width", type=int, required=True)
    p.add_argument("--max_disp", type=int, default=192)
    p.add_argument("--resume", type=str, default="")
    p.add_argument("--left_right", type=int, default=0)
    p.add_argument("--batchSize", type=int, default=1, help="global batch size")
    p.add_argument("--testBatchSize", type=int, default=1)
    p.add_argument("--nEpochs", type=int, default=2048)
    p.add_argument("--lr", type=float, default=0.001)
    p.add_argument("--cuda", type=int, default=1)
    p.add_argument("--threads", type=int, default=1)
    p.add_argument("--seed", type=int, default=123)
    p.add_argument("--shift", type=int, default=0)
    p.add_argument("--kitti", type=int, default=0)
    p.add_argument("--kitti2015", type=int, default=0)
    p.add_argument("--data_path", type=str, default="/ssd1/zhangfeihu/data/stereo/")
    p.add_argument("--training_list", type=str, default="./lists/sceneflow_train.list")
    p.add_argument("--val_list", type=str, default="./lists/sceneflow_test_select.list")
    p.add_argument("--save_path", type=str, default="./checkpoint/")
    p.add_argument("--model", type=str, default="GANet_deep")
    # additions
    p.add_argument("--synthetic", type=int, default=0, help="train on this many generated pairs per epoch")
    p.add_argument("--models_dir", type=str, default="", help="directory holding the reference's models/*.py")
    p.add_argument("--max_iters", type=int, default=0, help="stop after this many iterations (smoke runs)")
    p.add_argument("--fuse_sga_blocks", type=int, default=1, help="fused SGABlock prologue and DispAgg tail (ganet_b200.fused)")
    return p.parse_args(argv)


def build_model(opt, device):
    from baseline import refmodels
    if opt.models_dir:
        refmodels.MODELS = opt.models_dir
    if opt.model not in ("GANet11", "GANet_deep"):
        raise Exception("No suitable model found ...")                # train.py:51
    model = refmodels.build(opt.model, opt.max_disp, seed=opt.seed, device=device)
    if opt.fuse_sga_blocks:
        from ganet_b200.fused import fuse_disp_heads, fuse_sga_blocks
        fuse_sga_blocks(model)
        fuse_disp_heads(model)
    return model


def load_checkpoint_into(model, path):
    """Accepts the reference's DataParallel checkpoints (`module.` keys) and bare ones (train.py:75-82)."""
    ck = torch.load(path, map_location="cpu")
    state = ck["state_dict"] if "state_dict" in ck else ck
    bare = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state.items()}
    target = model.module if hasattr(model, "module") else model
    missing, unexpected = target.load_state_dict(bare, strict=False)
    return ck.get("epoch", 0), missing, unexpected


def checkpoint_state(model, optimizer, epoch):
    """The reference's dictionary (train.py:196-200); keys prefixed `module.` like its DataParallel models'."""
    target = model.module if hasattr(model, "module") else model
    return {"epoch": epoch, "state_dict": {"module." + k: v for k, v in target.state_dict().items()},
            "optimizer": optimizer.state_dict()}


def train_loss(opt, outs, target, mask, criterion):
    """train.py:104-118."""
    if opt.model == "GANet11":
        d1, d2 = outs
        last = criterion(d2[mask], target[mask]) if (opt.kitti or opt.kitti2015) else F.smooth_l1_loss(d2[mask], target[mask])
        return 0.4 * F.smooth_l1_loss(d1[mask], target[mask]) + 1.2 * last, ((d1 + d2) / 2.0, d1, d2)
    d0, d1, d2 = outs
    last = criterion(d2[mask], target[mask]) if (opt.kitti or opt.kitti2015) else F.smooth_l1_loss(d2[mask], target[mask])
    return 0.2 * F.smooth_l1_loss(d0[mask], target[mask]) + 0.6 * F.smooth_l1_loss(d1[mask], target[mask]) + last, (d0, d1, d2)


def main(argv=None):
    opt = parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("harness/train_ddp.py: no CUDA device; the operators have no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(opt.seed)
    from harness.data import ListStereo, SyntheticStereo
    from libs.GANet.modules.GANet import MyLoss2

    if opt.synthetic:
        train_set = SyntheticStereo(opt.synthetic, opt.crop_height, opt.crop_width, opt.max_disp, seed=opt.seed)
    else:
        train_set = ListStereo(opt.data_path, opt.training_list, (opt.crop_height, opt.crop_width), True,
                               bool(opt.kitti), bool(opt.kitti2015), seed=opt.seed + rank)
    if opt.batchSize % world:
        raise SystemExit("--batchSize must be a multiple of the number of GPUs")
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, world, rank, shuffle=True, drop_last=True) if world > 1 else None
    loader = torch.utils.data.DataLoader(train_set, batch_size=opt.batchSize // world, shuffle=sampler is None,
                                         sampler=sampler, num_workers=opt.threads, drop_last=True, pin_memory=True)

    model = build_model(opt, dev)
    criterion = MyLoss2(thresh=3, alpha=2)                              # train.py:71
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    optimizer = torch.optim.Adam(model.parameters(), lr=opt.lr, betas=(0.9, 0.999))
    if opt.resume:
        if os.path.isfile(opt.resume):
            ep, missing, unexpected = load_checkpoint_into(model, opt.resume)
            if rank == 0:
                print("=> loaded checkpoint '%s' (epoch %s, %d missing, %d unexpected keys)" % (opt.resume, ep, len(missing), len(unexpected)))
        elif rank == 0:
            print("=> no checkpoint found at '%s'" % opt.resume)

    iters = 0
    for epoch in range(1, opt.nEpochs + 1):
        lr = opt.lr if epoch <= 400 else opt.lr * 0.1                   # train.py:171-178
        for group in optimizer.param_groups:
            group["lr"] = lr
        if sampler is not None:
            sampler.set_epoch(epoch)
        model.train()
        sums, n_ok, t0 = [0.0] * 4, 0, time.time()
        for it, (left, right, target) in enumerate(loader):
            left, right = left.to(dev, non_blocking=True), right.to(dev, non_blocking=True)
            target = target.to(dev, non_blocking=True).squeeze(1)
            mask = target < opt.max_disp
            if int(mask.sum()) > 0 or world > 1:                        # under DDP every rank must step together
                optimizer.zero_grad(set_to_none=True)
                loss, (d0, d1, d2) = train_loss(opt, model(left, right), target, mask, criterion)
                loss.backward()
                optimizer.step()
                errs = [float(torch.mean(torch.abs(d.detach()[mask] - target[mask]))) for d in (d0, d1, d2)]
                sums = [sums[0] + float(loss.detach())] + [s + e for s, e in zip(sums[1:], errs)]
                n_ok += 1
                if rank == 0:
                    print("===> Epoch[{}]({}/{}): Loss: {:.4f}, Error: ({:.4f} {:.4f} {:.4f})".format(
                        epoch, it, len(loader), float(loss.detach()), *errs))
                    sys.stdout.flush()
            iters += 1
            if opt.max_iters and iters >= opt.max_iters:
                break
        if rank == 0 and n_ok:
            print("===> Epoch {} Complete: Avg. Loss: {:.4f}, Avg. Error: ({:.4f} {:.4f} {:.4f}) [{:.1f}s]".format(
                epoch, *[s / n_ok for s in sums], time.time() - t0))
        save_now = (epoch % 50 == 0 and epoch >= 300) if (opt.kitti or opt.kitti2015) else epoch >= 8     # :191-205
        last = epoch == opt.nEpochs or (opt.max_iters and iters >= opt.max_iters)
        if rank == 0 and (save_now or last):
            os.makedirs(os.path.dirname(opt.save_path) or ".", exist_ok=True)
            name = opt.save_path + "_epoch_{}.pth".format(epoch)
            torch.save(checkpoint_state(model, optimizer, epoch), name)
            print("Checkpoint saved to {}".format(name))
        if last:
            break
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
