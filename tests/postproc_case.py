"""Seeded inputs of the FCOS post-processor parity cases (shared by tests/tools/make_golden_postproc.py, the oracle test and the
GPU test): per-level class probability maps, exp()-ed regression maps, centerness logits."""
import numpy as np
import torch

SHAPES = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]
STRIDES = [8, 16, 32, 64, 128]
# name -> (images, fg classes, seed, fraction of confident entries, pre_nms_thresh, pre_nms_top_n, nms_thresh, post_top_n)
CASES = {
    "c8": (2, 8, 31, 0.003, 0.05, 1000, 0.6, 100),        # below every cap: the reference order is well defined
    "topk": (2, 8, 32, 0.08, 0.05, 60, 0.6, 100),       # levels with more candidates than pre_nms_top_n; cap at 100 active
    "car_cap": (3, 1, 33, 0.15, 0.05, 1000, 0.5, 20),   # single class, heavy overlap, detections-per-image cap of 20
}
IMAGE_SIZES = [(200, 336), (190, 300), (200, 336)]


def build(name):
    n, c, seed, frac, thr, top_n, nms_thr, post_n = CASES[name]
    rs = np.random.RandomState(seed)
    probs, regs, ctrs = [], [], []
    for (h, w), s in zip(SHAPES, STRIDES):
        p = rs.uniform(0.0, 0.04, (n, c, h, w))
        hot = rs.rand(n, c, h, w) < frac
        p[hot] = rs.uniform(0.06, 0.99, int(hot.sum()))
        probs.append(torch.from_numpy(p.astype(np.float32)))
        regs.append(torch.from_numpy(np.exp(rs.standard_normal((n, 4, h, w)) * 0.5 + np.log(s * 2.0)).astype(np.float32)))
        ctrs.append(torch.from_numpy(rs.standard_normal((n, 1, h, w)).astype(np.float32)))
    return dict(shapes=SHAPES, strides=STRIDES, probs=probs, regs=regs, ctrs=ctrs, sizes=IMAGE_SIZES[:n], thr=thr, top_n=top_n,
                nms_thr=nms_thr, post_n=post_n, num_fg=c)


def canonical(boxes, scores, labels):
    """Order-independent form of one image's detections: rows (label, score, x1, y1, x2, y2) sorted lexicographically."""
    a = np.concatenate([np.asarray(labels, dtype=np.float64)[:, None], np.asarray(scores, dtype=np.float64)[:, None],
                        np.asarray(boxes, dtype=np.float64).reshape(-1, 4)], axis=1)
    if a.shape[0] == 0:
        return a
    return a[np.lexsort(a.T[::-1])]
