"""Generate tests/golden/fcos_loss.npz from the UNMODIFIED reference FCOSLossComputation (build container only).

    PYTHONPATH=. python tests/tools/make_golden_fcos_loss.py

Seeded head outputs (classification logits, exp()-ed regression maps, centerness logits) at the small test geometry, ragged
box lists; the reference runs on the CPU (its SigmoidFocalLoss then takes `sigmoid_focal_loss_cpu`, which indexes gamma[0] /
alpha[0]: the cfg values are passed as 1-element lists).  Stored: the three losses and their gradients w.r.t. all 15 maps.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
from oracle import condgraph_oracle as orc  # noqa: E402
import fcos_loss_case  # noqa: E402


def main():
    _, loss_mod, BoxList = ref_shim.reference_modules()
    cfg = ref_shim.to_cfgnode({"MODEL": {"FCOS": {"LOSS_GAMMA": [2.0], "LOSS_ALPHA": [0.25]}}})
    ref = loss_mod.FCOSLossComputation(cfg)
    out = {}
    for name in fcos_loss_case.CASES:
        shapes, strides, boxes, labels, cls, reg, ctr, hw = fcos_loss_case.build(name)
        targets = []
        for b, l in zip(boxes, labels):
            t = BoxList(b, (hw[1], hw[0]), mode="xyxy")
            t.add_field("labels", l)
            targets.append(t)
        locations = [torch.stack(orc.level_locations(h, w, s), dim=1) for (h, w), s in zip(shapes, strides)]
        maps = [m.clone().requires_grad_(True) for m in cls + reg + ctr]
        n = len(shapes)
        losses = ref(locations, maps[:n], maps[n:2 * n], maps[2 * n:], targets)
        (losses[0] * 1.0 + losses[1] * 0.7 + losses[2] * 1.3).backward()
        for i, v in enumerate(losses):
            out["%s/loss%d" % (name, i)] = np.array(float(v), dtype=np.float64)
        for i, m in enumerate(maps):
            out["%s/grad%d" % (name, i)] = m.grad.numpy().copy()
        print(name, [float(v) for v in losses])
    meta = "reference=/root/reference FCOSLossComputation torch=%s" % torch.__version__
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fcos_loss.npz"), __meta__=np.array(meta), **out)


if __name__ == "__main__":
    main()
