"""Stereo pairs for the harness.

`SyntheticStereo` needs no files: a random texture is warped by a smooth random disparity field, so a network
can actually fit it (the offline container has neither SceneFlow nor KITTI).  `ListStereo` reads the
reference's list files (lists/*.list) with the directory conventions of dataloader/dataset.py -- SceneFlow
(frames_finalpass/ + disparity/*.pfm, :118-131), KITTI 2012 (colored_0/, colored_1/, disp_occ/, :150-160),
KITTI 2015 (image_2/, image_3/, disp_occ_0/, :194-204) -- with the same per-channel standardisation
(:133-144), invalid-disparity convention (KITTI: 0 -> huge, :168-170) and crop rules (random crop in training,
bottom-right-aligned zero padding or centre crop in testing, :93-111).  Needs only numpy + PIL (the reference
needs skimage, which this image does not have)."""
import re

import numpy as np
import torch
from torch.utils import data


def read_pfm(path):
    """Portable float map -> (H, W) float32, top row first (dataloader/dataset.py:12-46 returns the same)."""
    with open(path, "rb") as fh:
        kind = fh.readline().decode("latin-1").strip()
        if kind not in ("PF", "Pf"):
            raise ValueError("%s: not a PFM file" % path)
        channels = 3 if kind == "PF" else 1
        w, h = (int(v) for v in re.findall(r"\d+", fh.readline().decode("latin-1")))
        scale = float(fh.readline().decode("latin-1").strip())
        img = np.frombuffer(fh.read(w * h * channels * 4), dtype="<f4" if scale < 0 else ">f4")
    img = img.reshape(h, w, channels)[::-1, :, 0] if channels == 3 else img.reshape(h, w)[::-1]
    return np.ascontiguousarray(img, dtype=np.float32)


def standardise(img):
    """(H, W, 3) uint8 -> (3, H, W) float32, zero mean / unit std per channel (dataset.py:133-144)."""
    a = np.asarray(img, dtype=np.float32)[:, :, :3].transpose(2, 0, 1)
    m = a.reshape(3, -1).mean(1)[:, None, None]
    s = a.reshape(3, -1).std(1)[:, None, None]
    return (a - m) / np.maximum(s, 1e-6)


def crop_or_pad(left, right, disp, ch, cw, rng=None):
    """Training (rng given): random crop, zero-padding first if the image is smaller (invalid disparity = 1000,
    dataset.py:52-64).  Testing: pad at the top-left so the image sits bottom-right (:99-103), else centre crop."""
    _, h, w = left.shape
    if h < ch or w < cw:
        ph, pw = max(ch, h), max(cw, w)
        pad = lambda a, fill: np.pad(a, ((0, 0), (ph - h, 0), (pw - w, 0)), constant_values=fill)   # noqa: E731
        left, right, disp = pad(left, 0.0), pad(right, 0.0), pad(disp, 1000.0)
        h, w = ph, pw
    if rng is not None:
        y0 = int(rng.integers(0, h - ch + 1)); x0 = int(rng.integers(0, w - cw + 1))
    else:
        y0, x0 = (h - ch) // 2, (w - cw) // 2
    sl = (slice(None), slice(y0, y0 + ch), slice(x0, x0 + cw))
    return left[sl], right[sl], disp[sl]


class SyntheticStereo(data.Dataset):
    def __init__(self, n, crop_height, crop_width, max_disp, seed=0):
        self.n, self.h, self.w, self.max_disp, self.seed = n, crop_height, crop_width, max_disp, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        h, w = self.h, self.w
        tex = torch.nn.functional.interpolate(torch.randn(1, 3, h // 4 + 1, w // 4 + 1, generator=g), size=(h, w),
                                              mode="bilinear", align_corners=False)
        left = tex + 0.3 * torch.randn(1, 3, h, w, generator=g)
        low = torch.rand(1, 1, 3, 4, generator=g) * min(self.max_disp - 1, w // 4)
        disp = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True)
        xs = torch.arange(w).view(1, 1, 1, w).float() + disp          # right(x - d) = left(x): sample left at x + d
        grid = torch.stack(((xs / (w - 1)) * 2 - 1, (torch.arange(h).view(1, 1, h, 1).float() / max(h - 1, 1) * 2 - 1).expand(1, 1, h, w)), -1)[:, 0]
        right = torch.nn.functional.grid_sample(left, grid, mode="bilinear", padding_mode="border", align_corners=True)
        norm = lambda t: (t - t.mean((2, 3), keepdim=True)) / t.std((2, 3), keepdim=True)   # noqa: E731
        return norm(left)[0], norm(right)[0], disp[0]


class ListStereo(data.Dataset):
    def __init__(self, data_path, file_list, crop_size, training=True, kitti=False, kitti2015=False, seed=0):
        from PIL import Image
        self._open = Image.open
        self.root = data_path
        with open(file_list) as fh:
            self.files = [ln.strip() for ln in fh if ln.strip()]
        self.ch, self.cw = crop_size
        self.training, self.kitti, self.kitti2015 = training, kitti, kitti2015
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        return len(self.files)

    def _load(self, name):
        r = self.root
        if self.kitti or self.kitti2015:
            dirs = ("colored_0/", "colored_1/", "disp_occ/") if self.kitti else ("image_2/", "image_3/", "disp_occ_0/")
            left, right = self._open(r + dirs[0] + name), self._open(r + dirs[1] + name)
            d = np.asarray(self._open(r + dirs[2] + name), dtype=np.float32)
            w = d.shape[1]
            d = np.where(d < 0.1, w * 2 * 256.0, d) / 256.0           # dataset.py:168-170
        else:
            left = self._open(r + "frames_finalpass/" + name)
            right = self._open(r + "frames_finalpass/" + name[:-13] + "right/" + name[-8:])
            d = read_pfm(r + "disparity/" + name[:-3] + "pfm")
        return standardise(left), standardise(right), d[None].astype(np.float32)

    def __getitem__(self, i):
        left, right, disp = self._load(self.files[i])
        left, right, disp = crop_or_pad(left, right, disp, self.ch, self.cw, self.rng if self.training else None)
        return (torch.from_numpy(np.ascontiguousarray(left)), torch.from_numpy(np.ascontiguousarray(right)),
                torch.from_numpy(np.ascontiguousarray(disp)))
