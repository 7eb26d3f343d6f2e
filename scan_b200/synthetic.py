"""Seeded synthetic workload of Cityscapes shape.

FPN features: 5 fp32 NCHW tensors [N,256,H_l,W_l], i.i.d. N(0,1) plus, inside every GT box,
a fixed unit-norm per-class direction scaled by 2 (gives class structure to attention,
DBSCAN and the softmax).  Targets: `boxes_per_image` boxes, side exp(U(log16,log400)),
aspect U(0.4,2.5), uniformly placed in the padded image, labels U{1..num_fg}.
Everything is generated on the CPU from `numpy.random.RandomState` (MT19937, a stream that is
stable across numpy/torch versions) so that the oracle (CPU), the CUDA path and the committed
golden fixtures all see bit-identical inputs.
"""
import math
import numpy as np
import torch

from .config import CITYSCAPES_LEVEL_SHAPES
from .structures import BoxList


def level_shapes_for(image_hw=(800, 1344), strides=(8, 16, 32, 64, 128)):
    h, w = image_hw
    return [(int(math.ceil(h / s)), int(math.ceil(w / s))) for s in strides]


def make_boxes(n_images, num_fg, boxes_per_image=18, image_hw=(800, 1344), seed=1234):
    rs = np.random.RandomState(seed)
    H, W = image_hw
    out = []
    for _ in range(n_images):
        side = np.exp(rs.uniform(math.log(16.0), math.log(400.0), boxes_per_image))
        aspect = rs.uniform(0.4, 2.5, boxes_per_image)
        bw = np.minimum(side * np.sqrt(aspect), W - 2.0)
        bh = np.minimum(side / np.sqrt(aspect), H - 2.0)
        x0 = rs.uniform(0.0, 1.0, boxes_per_image) * (W - 1.0 - bw)
        y0 = rs.uniform(0.0, 1.0, boxes_per_image) * (H - 1.0 - bh)
        boxes = torch.from_numpy(np.stack([x0, y0, x0 + bw, y0 + bh], axis=1).astype(np.float32))
        labels = torch.from_numpy(rs.randint(1, num_fg + 1, boxes_per_image).astype(np.int64))
        out.append((boxes, labels))
    return out


def make_features(n_images, num_fg, level_shapes=None, strides=(8, 16, 32, 64, 128), channels=256,
                  boxes=None, seed=1234, signal=2.0, dir_seed=None):
    """Returns list of 5 fp32 NCHW CPU tensors.  `dir_seed` fixes the per-class directions independently of the
    noise seed (source and target domains of the benchmark share the classes but not the images)."""
    if level_shapes is None:
        level_shapes = CITYSCAPES_LEVEL_SHAPES
    rs = np.random.RandomState(seed + 7919)
    drs = rs if dir_seed is None else np.random.RandomState(dir_seed)
    dirs = torch.from_numpy(drs.standard_normal((num_fg + 1, channels)).astype(np.float32))
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    feats = []
    for (h, w), s in zip(level_shapes, strides):
        f = torch.from_numpy(rs.standard_normal((n_images, channels, h, w)).astype(np.float32))
        if boxes is not None:
            ys = (torch.arange(h, dtype=torch.float32) * s + s // 2)[:, None]
            xs = (torch.arange(w, dtype=torch.float32) * s + s // 2)[None, :]
            for n, (bx, lb) in enumerate(boxes):
                for k in range(bx.shape[0]):
                    inside = (xs > bx[k, 0]) & (xs < bx[k, 2]) & (ys > bx[k, 1]) & (ys < bx[k, 3])
                    f[n] += signal * dirs[int(lb[k])][:, None, None] * inside[None].float()
        feats.append(f.contiguous())
    return feats


def make_workload(n_images, num_fg, seed=1234, boxes_per_image=18, level_shapes=None, image_hw=(800, 1344),
                  boxlist_cls=BoxList, with_targets=True, signal=2.0, dir_seed=None):
    boxes = make_boxes(n_images, num_fg, boxes_per_image, image_hw, seed)
    feats = make_features(n_images, num_fg, level_shapes=level_shapes, boxes=boxes, seed=seed, signal=signal,
                          dir_seed=dir_seed)
    targets = None
    if with_targets:
        targets = []
        for bx, lb in boxes:
            t = boxlist_cls(bx, (image_hw[1], image_hw[0]), mode="xyxy")
            t.add_field("labels", lb)
            targets.append(t)
    return feats, targets
