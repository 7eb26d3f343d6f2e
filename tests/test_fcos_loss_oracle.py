"""f2: the oracle restatement of FCOSLossComputation against the golden vectors generated from the unmodified reference
(tests/tools/make_golden_fcos_loss.py), losses and all 15 map gradients."""
import os

import numpy as np
import pytest
import torch

import fcos_loss_case
from oracle import condgraph_oracle as orc


@pytest.mark.parametrize("name", list(fcos_loss_case.CASES))
def test_oracle_fcos_loss_matches_reference_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, "fcos_loss.npz"))
    shapes, strides, boxes, labels, cls, reg, ctr, _ = fcos_loss_case.build(name)
    maps = [m.clone().requires_grad_(True) for m in cls + reg + ctr]
    n = len(shapes)
    losses = orc.fcos_loss_computation(shapes, strides, boxes, labels, maps[:n], maps[n:2 * n], maps[2 * n:], 2.0, 0.25)
    (losses[0] * 1.0 + losses[1] * 0.7 + losses[2] * 1.3).backward()
    for i, v in enumerate(losses):
        assert abs(float(v) - float(gold["%s/loss%d" % (name, i)])) <= 1e-5 * max(1.0, abs(float(v)))
    for i, m in enumerate(maps):
        w = gold["%s/grad%d" % (name, i)]
        assert np.abs(m.grad.numpy() - w).max() <= 1e-5 * max(np.abs(w).max(), 1e-12) + 1e-9, "grad %d" % i
