"""Module-level parity (B200): scan_b200.GRAPHModule against the golden fixtures generated from the unmodified
reference, and against the CPU oracle at the full Cityscapes size."""
import os

import numpy as np
import pytest
import torch

import harness

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(harness.CASES))
def test_product_matches_reference_golden(name, golden_dir):
    from scan_b200.condgraph import build_condgraph
    want = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    want.pop("__meta__")
    cfg = harness.build_case(name)[0]
    got = harness.run_case(name, build_condgraph(cfg, 256), "product", device="cuda")
    bad = harness.compare(got, want, rtol=1e-3, device_run=True)
    assert not bad, "\n".join(bad[:25])


def test_full_size_source_target_eval_vs_oracle():
    """N=1, 800x1344 (22 400 locations), K=9: one source step, one target step, one eval -- live oracle."""
    from scan_b200.condgraph import build_condgraph
    from oracle.condgraph_oracle import build_oracle
    harness.CASES["_full"] = dict(cfg=("c2f", {}), n=1, steps=["source", "target", "eval"], fixture=(3.0, 2.0), full=True)
    cfg = harness.build_case("_full")[0]
    want = harness.run_case("_full", build_oracle(cfg), "oracle")
    got = harness.run_case("_full", build_condgraph(cfg, 256), "product", device="cuda")
    bad = harness.compare(got, want, rtol=1e-3, device_run=True)
    assert not bad, "\n".join(bad[:25])


def test_double_backward_call_on_source_graph():
    """engine/trainer.py:299,343 back-propagates the same source graph twice (retain_graph=True)."""
    from scan_b200.condgraph import build_condgraph
    cfg, case, src_feats, src_targets, _ = harness.build_case("c2f_small")
    m = build_condgraph(cfg, 256).cuda().train()
    m.multihead_attn.p_drop = 0.0
    feats = [f.cuda().requires_grad_(True) for f in src_feats]
    out, (node_loss, _), act_loss, acts = m(None, feats, targets=src_targets, mode="source")
    (node_loss + act_loss).backward(retain_graph=True)
    g1 = feats[0].grad.clone()
    sum((a * a).sum() for a in acts).backward()
    assert torch.isfinite(feats[0].grad).all() and not torch.equal(feats[0].grad, g1)
