"""the train / predict harness.  CPU: data pipeline, checkpoint key compatibility, the
reference's command-line flags.  GPU: a two-iteration training run of the reference's GANet-11 on generated pairs,
checkpoint in the reference's format, resume, and inference on a PNG pair."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_data_pipeline(tmp_path):
    from harness.data import ListStereo, SyntheticStereo, crop_or_pad, read_pfm, standardise
    l, r, d = SyntheticStereo(3, 48, 96, 48, seed=1)[2]
    assert l.shape == (3, 48, 96) and r.shape == (3, 48, 96) and d.shape == (1, 48, 96)
    assert abs(float(l.mean())) < 1e-5 and abs(float(l.std()) - 1) < 1e-3 and 0 <= float(d.min()) <= float(d.max()) < 48
    img = np.random.default_rng(0).random((4, 6)).astype("<f4")
    p = tmp_path / "d.pfm"
    p.write_bytes(b"Pf\n6 4\n-1.0\n" + img[::-1].tobytes())           # PFM stores the bottom row first
    assert np.array_equal(read_pfm(str(p)), img)
    a = np.arange(3 * 5 * 7, dtype=np.float32).reshape(3, 5, 7)
    L, R, D = crop_or_pad(a, a, a[:1], 8, 8)                            # test mode: padded, image bottom-right
    assert L.shape == (3, 8, 8) and D[0, 0, 0] == 1000.0 and L[0, -1, -1] == a[0, -1, -1] and L[0, 0, 0] == 0
    L, _, _ = crop_or_pad(a, a, a[:1], 3, 5)                            # centre crop
    assert np.array_equal(L, a[:, 1:4, 1:6])
    s = standardise((np.random.default_rng(1).random((5, 6, 3)) * 255).astype(np.uint8))
    assert s.shape == (3, 5, 6) and np.allclose(s.reshape(3, -1).mean(1), 0, atol=1e-5)
    # a two-frame KITTI-2015-style tree read through ListStereo
    from PIL import Image
    for sub in ("image_2", "image_3", "disp_occ_0"):
        os.makedirs(tmp_path / sub)
    rng = np.random.default_rng(2)
    for name in ("000000_10.png", "000001_10.png"):
        for sub in ("image_2", "image_3"):
            Image.fromarray((rng.random((20, 30, 3)) * 255).astype(np.uint8)).save(tmp_path / sub / name)
        disp = (rng.random((20, 30)) * 40 * 256).astype(np.uint16)
        disp[0, 0] = 0                                                  # invalid pixel
        Image.fromarray(disp).save(tmp_path / "disp_occ_0" / name)
    lst = tmp_path / "val.list"
    lst.write_text("000000_10.png\n000001_10.png\n")
    ds = ListStereo(str(tmp_path) + "/", str(lst), (16, 24), training=True, kitti2015=True, seed=3)
    left, right, d = ds[1]
    assert len(ds) == 2 and left.shape == (3, 16, 24) and d.shape == (1, 16, 24)
    full = ListStereo(str(tmp_path) + "/", str(lst), (48, 48), training=False, kitti2015=True)[0][2]
    assert full.shape == (1, 48, 48) and float(full[0, 0, 0]) == 1000.0 and float(full.max()) >= 30 * 2      # padding / invalid


def test_checkpoint_keys_and_flags(tmp_path):
    from harness.train_ddp import checkpoint_state, load_checkpoint_into, parse_args
    net = torch.nn.Sequential(torch.nn.Conv2d(2, 3, 1), torch.nn.BatchNorm2d(3))
    opt = torch.optim.Adam(net.parameters())
    ck = checkpoint_state(net, opt, 7)
    assert set(ck) == {"epoch", "state_dict", "optimizer"} and all(k.startswith("module.") for k in ck["state_dict"])
    p = tmp_path / "ck_epoch_7.pth"
    torch.save(ck, p)
    other = torch.nn.Sequential(torch.nn.Conv2d(2, 3, 1), torch.nn.BatchNorm2d(3))
    ep, missing, unexpected = load_checkpoint_into(other, str(p))              # DataParallel-style keys
    assert ep == 7 and not missing and not unexpected
    assert torch.equal(other[0].weight, net[0].weight)
    torch.save({"state_dict": net.state_dict()}, p)                            # bare keys load too
    assert load_checkpoint_into(other, str(p))[1] == []
    # the reference's train.sh command line parses unchanged (train.sh:1-11)
    o = parse_args("--batchSize=16 --crop_height=240 --crop_width=528 --max_disp=192 --thread=16 --data_path=/d/ "
                   "--training_list=lists/sceneflow_train.list --save_path=./checkpoint/sceneflow --resume= "
                   "--model=GANet_deep --nEpochs=11".split())
    assert (o.batchSize, o.crop_width, o.threads, o.model, o.nEpochs) == (16, 528, 16, "GANet_deep", 11)


@pytest.mark.gpu
def test_train_resume_predict_round_trip(tmp_path):
    from baseline import refmodels
    if not refmodels.available():
        pytest.skip("reference models not copied")
    from PIL import Image
    from harness import infer as predict, train_ddp
    save = str(tmp_path / "ck")
    args = ["--crop_height", "48", "--crop_width", "96", "--max_disp", "48", "--model", "GANet11", "--synthetic", "4",
            "--batchSize", "1", "--nEpochs", "1", "--max_iters", "2", "--threads", "0", "--save_path", save]
    assert train_ddp.main(args) == 0
    ck = torch.load(save + "_epoch_1.pth", map_location="cpu")
    assert ck["epoch"] == 1 and len(ck["state_dict"]) == 364 and all(k.startswith("module.") for k in ck["state_dict"])
    assert all(torch.isfinite(v).all() for v in ck["state_dict"].values() if v.is_floating_point())
    assert train_ddp.main(args + ["--resume", save + "_epoch_1.pth"]) == 0        # resume from the reference's format
    rng = np.random.default_rng(0)
    for n in ("l.png", "r.png"):
        Image.fromarray((rng.random((40, 90, 3)) * 255).astype(np.uint8)).save(tmp_path / n)
    out = str(tmp_path / "out" / "disp.png")
    assert predict.main(["--crop_height", "48", "--crop_width", "96", "--max_disp", "48", "--model", "GANet11",
                         "--resume", save + "_epoch_1.pth", "--left", str(tmp_path / "l.png"),
                         "--right", str(tmp_path / "r.png"), "--save", out]) == 0
    d = np.asarray(Image.open(out))
    assert d.shape == (40, 90) and d.dtype == np.uint16                          # cropped back, disparity * 256
