"""Training-time augmentation on the device: the reference's ``RandomHorizontalFlip`` followed
by ``ScaleNRotate(rots=(-30, 30), scales=(.75, 1.25))`` (dataloaders/custom_transforms.py:7-54, :87-100; composed at
train_online.py:92-94 and train_parent.py:108-110), applied to a GPU-resident batch by one gather kernel per tensor
instead of cv2 on a DataLoader worker.  The random draws use Python's ``random`` in the reference's order (flip
first, then rotation, then scale) so a seeded run picks the same transformations.  Matrices follow OpenCV's
``getRotationMatrix2D`` / ``warpAffine`` (inverse map); the arithmetic is in csrc/augment.cu.
"""
import ctypes
import math
import random

import torch

from . import _native as nat
from . import ops


def rotation_matrix(center, angle_deg, scale):
    """cv2.getRotationMatrix2D: [[a, b, (1-a)cx - b cy], [-b, a, b cx + (1-a) cy]], a = s cos, b = s sin."""
    a = scale * math.cos(math.radians(angle_deg))
    b = scale * math.sin(math.radians(angle_deg))
    cx, cy = center
    return [a, b, (1.0 - a) * cx - b * cy, -b, a, b * cx + (1.0 - a) * cy]


def invert_affine(m):
    """The dst->src matrix cv::warpAffine derives from M (no WARP_INVERSE_MAP)."""
    m = list(m)
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0], m[1], m[3], m[4] = a11, m[1] * -d, m[3] * -d, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def draw_params(n, rots=(-30, 30), scales=(.75, 1.25), rng=random):
    """Per-sample (flip, rot, scale) drawn like the reference's transforms: RandomHorizontalFlip.__call__ draws
    first (custom_transforms.py:92), then ScaleNRotate draws rot and sc (:25-29)."""
    out = []
    for _ in range(n):
        flip = rng.random() < 0.5
        rot = (rots[1] - rots[0]) * rng.random() - (rots[1] - rots[0]) / 2
        sc = (scales[1] - scales[0]) * rng.random() - (scales[1] - scales[0]) / 2 + 1
        out.append((flip, rot, sc))
    return out


def affine_warp(x, params, mode):
    """x [n,c,h,w] fp32 CUDA -> warped copy.  params: list of n (flip, rot_degrees, scale); mode 'cubic'|'nearest'."""
    lib = nat.load()
    ops._require_cuda(x, "x")
    x = x.contiguous().float()
    n, c, h, w = (int(v) for v in x.shape)
    if len(params) != n:
        raise ValueError("one (flip, rot, scale) triple per sample")
    mats = (ctypes.c_double * (6 * n))()
    flips = (ctypes.c_int * n)()
    for i, (flip, rot, sc) in enumerate(params):
        inv = invert_affine(rotation_matrix((w / 2, h / 2), rot, sc))
        mats[6 * i:6 * i + 6] = inv
        flips[i] = int(bool(flip))
    out = torch.empty_like(x)
    ops._count((n + 31) // 32)
    with torch.cuda.device(x.device):
        nat.check(lib.osvos_affine_warp(x.data_ptr(), out.data_ptr(), mats, flips, n, c, h, w,
                                        0 if mode == "cubic" else 1, torch.cuda.current_stream().cuda_stream),
                  "osvos_affine_warp")
    return out


def augment_batch(sample, rots=(-30, 30), scales=(.75, 1.25), rng=random, params=None):
    """{'image': [n,3,h,w], 'gt': [n,1,h,w]} on the GPU -> augmented copy (image bicubic, gt nearest: DAVIS masks are
    0/1 after ``gt / gt.max()``, the case in which the reference selects INTER_NEAREST, custom_transforms.py:45-48)."""
    n = int(sample["image"].shape[0])
    params = draw_params(n, rots, scales, rng) if params is None else params
    return {"image": affine_warp(sample["image"], params, "cubic"), "gt": affine_warp(sample["gt"], params, "nearest")}
