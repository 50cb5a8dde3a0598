"""Golden fixtures for SURVEY.md 8(f) item 2 (augmentation), produced by the UNMODIFIED reference transforms.

Run in the build container only (needs /root/reference and cv2; neither is needed to USE the fixture):

    python tests/golden/make_golden_augment.py

It loads ``dataloaders/custom_transforms.py`` from /root/reference (no edits) and runs its own
``RandomHorizontalFlip`` followed by ``ScaleNRotate`` - i.e. ``cv2.flip``, ``cv2.getRotationMatrix2D`` and
``cv2.warpAffine`` with INTER_CUBIC for the image and INTER_NEAREST for the 0/1 mask - on seeded synthetic samples in
the layout the reference's DataLoader holds at that point (HWC fp32 image, HW fp32 mask).  Inputs (as uint8, the
values a decoded frame has), the random draws and the transforms' outputs are stored; ``tests/test_oracle.py`` pins
``oracle.osvos_oracle.scale_n_rotate`` to them and ``tests/test_gpu_augment.py`` the CUDA kernel.
"""
import importlib.util
import os
import random

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/dataloaders/custom_transforms.py"

# (height, width, python-random seed).  The draws cover both flip outcomes, both rotation signs and both sides of scale 1.
CASES = [(48, 70, 100), (48, 70, 101), (33, 45, 102), (33, 45, 104), (97, 131, 105), (97, 131, 106), (120, 214, 107)]
MEAN = (104.00699, 116.66877, 122.67892)     # dataloaders/davis_2016.py subtracts this before the transforms


def main():
    import cv2
    spec = importlib.util.spec_from_file_location("ref_custom_transforms", REF)
    ct = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ct)
    fx = {"cv2_version": cv2.__version__, "n_cases": len(CASES)}
    rng = np.random.default_rng(2024)
    for k, (h, w, seed) in enumerate(CASES):
        u8 = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        yy, xx = np.mgrid[0:h, 0:w]
        blob = ((yy - h * 0.45) ** 2 / (h * 0.3) ** 2 + (xx - w * 0.55) ** 2 / (w * 0.25) ** 2) < 1.0
        gt_u8 = (blob ^ (rng.random((h, w)) > 0.97)).astype(np.uint8)
        img = u8.astype(np.float32) - np.array(MEAN, dtype=np.float32)
        gt = gt_u8.astype(np.float32)
        random.seed(seed)
        sample = {"image": img.copy(), "gt": gt.copy(), "fname": "synthetic"}
        sample = ct.RandomHorizontalFlip()(sample)
        sample = ct.ScaleNRotate(rots=(-30, 30), scales=(.75, 1.25))(sample)
        random.seed(seed)                       # replay the draws in the transforms' order (:92, :25-29)
        flip = random.random() < 0.5
        rot = (30 - -30) * random.random() - (30 - -30) / 2
        sc = (1.25 - .75) * random.random() - (1.25 - .75) / 2 + 1
        fx[f"c{k}.image_u8"] = u8
        fx[f"c{k}.gt_u8"] = gt_u8
        fx[f"c{k}.draws"] = np.array([float(flip), rot, sc], dtype=np.float64)
        fx[f"c{k}.out_image"] = sample["image"].astype(np.float32)
        fx[f"c{k}.out_gt"] = sample["gt"].astype(np.float32)
        print(f"case {k}: {h}x{w} flip={flip} rot={rot:.3f} sc={sc:.4f} "
              f"mask pixels {int(sample['gt'].sum())}/{h * w}")
    path = os.path.join(HERE, "reference_augment.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
