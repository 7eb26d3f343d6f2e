"""Seeded inputs of the FCOSLossComputation parity cases (shared by tests/tools/make_golden_fcos_loss.py, the oracle test
and the GPU test)."""
import numpy as np
import torch

from scan_b200.synthetic import make_boxes

SHAPES = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]
STRIDES = [8, 16, 32, 64, 128]
HW = (200, 336)
# name -> (images, boxes per image, foreground classes, seed)
CASES = {"c8": (3, 6, 8, 11), "car": (2, 4, 1, 12), "dense": (2, 14, 8, 13)}


def build(name):
    n, nb, c, seed = CASES[name]
    rs = np.random.RandomState(seed)
    boxes = make_boxes(n, c, nb, HW, seed)
    bl = [(b[: max(1, nb - i)], l[: max(1, nb - i)]) for i, (b, l) in enumerate(boxes)]       # ragged counts
    cls = [torch.from_numpy(rs.standard_normal((n, c, h, w)).astype(np.float32) * 2 - 2) for h, w in SHAPES]
    reg = [torch.from_numpy(np.exp(rs.standard_normal((n, 4, h, w)) * 0.7 + np.log(s * 1.5)).astype(np.float32))
           for (h, w), s in zip(SHAPES, STRIDES)]
    ctr = [torch.from_numpy(rs.standard_normal((n, 1, h, w)).astype(np.float32)) for h, w in SHAPES]
    return SHAPES, STRIDES, [b for b, _ in bl], [l for _, l in bl], cls, reg, ctr, HW
