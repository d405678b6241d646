#!/usr/bin/env python
"""Inference harness: the reference's predict.py on the B200 operators.

    python harness/infer.py --crop_height 384 --crop_width 1248 --model GANet_deep --resume ckpt.pth \
        --left l.png --right r.png --save out.png            # one pair
    python harness/infer.py ... --kitti2015 1 --data_path D/ --test_list lists/kitti2015_val.list --save_path out/

Pre-processing (per-channel standardisation, padding so the image sits bottom-right, or centre crop) and the
output format (disparity * 256 as 16-bit PNG, cropped back to the image) follow predict.py:75-138; PIL writes the
PNG (the reference uses skimage, absent here).  One image per call, eval mode, no_grad."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="GANet inference on B200")
    p.add_argument("--crop_height", type=int, required=True)
    p.add_argument("--crop_width", type=int, required=True)
    p.add_argument("--max_disp", type=int, default=192)
    p.add_argument("--resume", type=str, default="")
    p.add_argument("--cuda", type=int, default=1)
    p.add_argument("--kitti", type=int, default=0)
    p.add_argument("--kitti2015", type=int, default=0)
    p.add_argument("--data_path", type=str, default="")
    p.add_argument("--test_list", type=str, default="")
    p.add_argument("--save_path", type=str, default="./result/")
    p.add_argument("--model", type=str, default="GANet_deep")
    p.add_argument("--left", type=str, default="")
    p.add_argument("--right", type=str, default="")
    p.add_argument("--save", type=str, default="")
    p.add_argument("--models_dir", type=str, default="")
    p.add_argument("--seed", type=int, default=123)
    return p.parse_args(argv)


def prepare(left_img, right_img, ch, cw):
    """-> left, right (1,3,ch,cw) tensors, original h, w (predict.py:75-114)."""
    from harness.data import standardise
    left, right = standardise(left_img), standardise(right_img)
    _, h, w = left.shape
    both = np.concatenate([left, right], 0)
    if h <= ch and w <= cw:
        canvas = np.zeros((6, ch, cw), np.float32)
        canvas[:, ch - h:, cw - w:] = both
    else:
        y0, x0 = int((h - ch) / 2), int((w - cw) / 2)
        canvas = both[:, y0:y0 + ch, x0:x0 + cw]
    t = torch.from_numpy(np.ascontiguousarray(canvas))[None]
    return t[:, :3].contiguous(), t[:, 3:].contiguous(), h, w


def predict_pair(model, left_img, right_img, ch, cw, device):
    left, right, h, w = prepare(left_img, right_img, ch, cw)
    model.eval()
    with torch.no_grad():
        disp = model(left.to(device), right.to(device))[0].float().cpu().numpy()
    if h <= ch and w <= cw:
        disp = disp[ch - h:, cw - w:]
    return disp


def save_disparity(path, disp):
    from PIL import Image
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    Image.fromarray((np.clip(disp, 0, 255.99) * 256).astype(np.uint16)).save(path)     # predict.py:138


def main(argv=None):
    opt = parse_args(argv)
    from PIL import Image
    from harness.train_ddp import build_model, load_checkpoint_into
    if not torch.cuda.is_available():
        raise SystemExit("harness/infer.py: no CUDA device; the operators have no CPU path")
    dev = torch.device("cuda", 0)
    opt.fuse_sga_blocks = 1
    model = build_model(opt, dev)
    if opt.resume:
        load_checkpoint_into(model, opt.resume)
    jobs = []
    if opt.left:
        jobs.append((opt.left, opt.right, opt.save or os.path.join(opt.save_path, os.path.basename(opt.left))))
    elif opt.test_list:
        dirs = ("colored_0/", "colored_1/") if opt.kitti else ("image_2/", "image_3/")
        for line in open(opt.test_list):
            name = line.strip()
            if name:
                jobs.append((opt.data_path + dirs[0] + name, opt.data_path + dirs[1] + name, opt.save_path + name))
    for lname, rname, out in jobs:
        disp = predict_pair(model, Image.open(lname), Image.open(rname), opt.crop_height, opt.crop_width, dev)
        save_disparity(out, disp)
        print("saved", out, disp.shape)
    return 0


if __name__ == "__main__":
    sys.exit(main())
