"""Diagnostic 2: gradients at the middle of the graph (d features_in, d act maps) product vs oracle, per level."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import harness
from oracle.condgraph_oracle import build_oracle
from scan_b200.condgraph import build_condgraph
from scan_b200.fixtures import fixture_state_dict

name = sys.argv[1] if len(sys.argv) > 1 else "c2f_small"
cfg, case, src_feats, src_targets, _ = harness.build_case(name)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False

def run(module, dev, variant):
    module.load_state_dict(fixture_state_dict(module, seed=99, kernel_gain=case["fixture"][0], gn_gain=case["fixture"][1]))
    module.to(dev); module.train(); module.multihead_attn.p_drop = 0.0
    feats = [f.clone().to(dev).requires_grad_(True) for f in src_feats]
    out_feats, (node_loss, _), act_loss, acts = module(None, feats, targets=src_targets, mode="source")
    fin = module.last["features_in"]
    for t in list(fin) + list(acts):
        t.retain_grad()
    gf = harness.cotangents([f.shape for f in out_feats], 1000); ga = harness.cotangents([a.shape for a in acts], 2000)
    total = 0
    if "feat" in variant: total = total + sum((f * g.to(dev)).sum() for f, g in zip(out_feats, gf))
    if "act" in variant: total = total + sum((a * g.to(dev)).sum() for a, g in zip(acts, ga))
    if "node" in variant: total = total + node_loss
    if "loss" in variant: total = total + act_loss
    total.backward()
    return ([t.grad.detach().cpu() if t.grad is not None else None for t in fin], [t.grad.detach().cpu() if t.grad is not None else None for t in acts],
            [f.grad.detach().cpu() for f in feats])

def rel(a, b):
    if a is None or b is None: return float("nan")
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))

for variant in ["feat", "act", "node", "loss", "feat+act+node+loss"]:
    o = run(build_oracle(cfg), "cpu", variant)
    p = run(build_condgraph(cfg, 256), "cuda", variant)
    print("variant %-20s" % variant)
    for nm, a, b in (("d_features_in", p[0], o[0]), ("d_acts", p[1], o[1]), ("d_input", p[2], o[2])):
        print("   %-14s relL2 per level:" % nm, ["%.1e" % rel(x, y) for x, y in zip(a, b)])
