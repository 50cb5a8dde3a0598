"""SURVEY.md 8(f) item 2 - device-side RandomHorizontalFlip + ScaleNRotate against the oracle's restatement of the
reference's cv2 path (dataloaders/custom_transforms.py:7-54, :87-100).  Coordinates are integer arithmetic and must
agree exactly (nearest mode is bit-exact); cubic values differ only by fp32 summation order / FMA contraction."""
import random

import numpy as np
import pytest
import torch

from oracle import osvos_oracle as oc

pytestmark = pytest.mark.gpu

CASES = [(0.0, 1.0, False), (0.0, 1.0, True), (180.0, 1.0, False), (17.0, 0.9, True), (-30.0, 1.25, False),
         (29.5, 0.75, True), (90.0, 1.0, False), (-3.3, 1.07, False)]


@pytest.mark.parametrize("shape", [(3, 40, 56), (3, 33, 45), (3, 120, 214)])
def test_affine_warp_matches_oracle(shape):
    from osvos_pytorch_b200 import augment
    g = torch.Generator().manual_seed(3)
    c, h, w = shape
    img = (torch.rand(len(CASES), c, h, w, generator=g) * 255.0 - 110.0)
    gt = (torch.rand(len(CASES), 1, h, w, generator=g) > 0.6).float()
    params = [(f, r, s) for (r, s, f) in CASES]
    out_i = augment.affine_warp(img.cuda(), params, "cubic").cpu().numpy()
    out_g = augment.affine_warp(gt.cuda(), params, "nearest").cpu().numpy()
    for k, (rot, sc, flip) in enumerate(CASES):
        want_i = oc.scale_n_rotate(img[k].numpy(), rot, sc, flip, nearest=False)
        want_g = oc.scale_n_rotate(gt[k].numpy(), rot, sc, flip, nearest=True)
        assert np.array_equal(out_g[k], want_g), (k, rot, sc, flip)                       # bit-exact
        err = np.abs(out_i[k] - want_i).max()
        assert err <= 2e-4, (k, rot, sc, flip, err)                                        # values up to ~255*1.3


def test_augment_batch_draws_like_the_reference_and_chunks():
    from osvos_pytorch_b200 import augment
    rng_a, rng_b = random.Random(7), random.Random(7)
    params = augment.draw_params(3, rng=rng_a)
    for flip, rot, sc in params:                       # same draw order as RandomHorizontalFlip then ScaleNRotate
        assert flip == (rng_b.random() < 0.5)
        assert rot == 60 * rng_b.random() - 30
        assert sc == 0.5 * rng_b.random() - 0.25 + 1
    n = 35                                             # > 32 samples: two kernel-parameter chunks
    g = torch.Generator().manual_seed(1)
    sample = {"image": torch.randn(n, 3, 24, 31, generator=g).cuda(), "gt": (torch.rand(n, 1, 24, 31, generator=g) > 0.5).float().cuda()}
    params = augment.draw_params(n, rng=random.Random(11))
    out = augment.augment_batch(sample, params=params)
    for k in (0, 31, 32, 34):
        flip, rot, sc = params[k]
        want = oc.scale_n_rotate(sample["image"][k].cpu().numpy(), rot, sc, flip, nearest=False)
        assert np.abs(out["image"][k].cpu().numpy() - want).max() < 1e-4
        want_g = oc.scale_n_rotate(sample["gt"][k].cpu().numpy(), rot, sc, flip, nearest=True)
        assert np.array_equal(out["gt"][k].cpu().numpy(), want_g)
    with pytest.raises(ValueError):
        augment.affine_warp(sample["image"], params[:3], "cubic")
    with pytest.raises(RuntimeError):
        augment.affine_warp(sample["image"].cpu(), params, "cubic")


def test_affine_warp_matches_the_reference_fixture():
    """The CUDA kernel against outputs of the reference's own cv2 transforms (tests/golden/reference_augment.npz,
    made by tests/golden/make_golden_augment.py from dataloaders/custom_transforms.py:7-54, :87-100): masks bit-exact,
    cubic pixels to fp32 summation order."""
    import os
    from osvos_pytorch_b200 import augment
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_augment.npz"))
    mean = np.array((104.00699, 116.66877, 122.67892), dtype=np.float32)
    for k in range(int(fx["n_cases"])):
        flip, rot, sc = fx[f"c{k}.draws"]
        img = torch.from_numpy((fx[f"c{k}.image_u8"].astype(np.float32) - mean).transpose(2, 0, 1).copy())[None]
        gt = torch.from_numpy(fx[f"c{k}.gt_u8"].astype(np.float32))[None, None]
        params = [(bool(flip), float(rot), float(sc))]
        out_i = augment.affine_warp(img.cuda(), params, "cubic").cpu().numpy()[0]
        out_g = augment.affine_warp(gt.cuda(), params, "nearest").cpu().numpy()[0, 0]
        assert np.array_equal(out_g, fx[f"c{k}.out_gt"]), k
        err = np.abs(out_i.transpose(1, 2, 0) - fx[f"c{k}.out_image"]).max()
        assert err <= 3e-4, (k, err)
