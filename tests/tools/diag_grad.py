"""Diagnostic: precision of torch's own CUDA ops (cuDNN conv fwd/bwd etc.) vs CPU under the tf32 flags, and
per-key product-vs-oracle errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))

g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(2, 256, 25, 42, generator=g)); w = torch.randn(256, 256, 3, 3, generator=g) * 0.01; b = torch.randn(256, generator=g) * 0.05
cot = torch.randn(2, 256, 25, 42, generator=g)
def conv_run(dev):
    xx, ww, bb = [t.clone().to(dev).requires_grad_(True) for t in (x, w, b)]
    y = F.conv2d(xx, ww, bb, padding=1)
    (y * cot.to(dev)).sum().backward()
    return y, xx.grad, ww.grad, bb.grad
ref = conv_run("cpu")
for setting in ["default", "legacy_off", "new_ieee", "deterministic"]:
    try:
        if setting == "legacy_off":
            torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        if setting == "new_ieee":
            torch.backends.cudnn.conv.fp32_precision = "ieee"; torch.backends.cuda.matmul.fp32_precision = "ieee"
        if setting == "deterministic":
            torch.backends.cudnn.deterministic = True; torch.backends.cudnn.benchmark = False
        out = conv_run("cuda")
        print("conv %-14s y %.2e dx %.2e dw %.2e db %.2e" % ((setting,) + tuple(rel(a, r) for a, r in zip(out, ref))))
    except Exception as e:
        print("conv", setting, "FAILED", repr(e)[:200])
# linear / layernorm / groupnorm backward
xl = torch.randn(600, 256, generator=g); wl = torch.randn(512, 256, generator=g) * 0.05; cl = torch.randn(600, 512, generator=g)
def lin(dev):
    a, ww = xl.clone().to(dev).requires_grad_(True), wl.clone().to(dev).requires_grad_(True)
    (F.linear(a, ww) * cl.to(dev)).sum().backward(); return a.grad, ww.grad
r = lin("cpu"); o = lin("cuda"); print("linear dx %.2e dw %.2e" % (rel(o[0], r[0]), rel(o[1], r[1])))
def gn(dev):
    a = x.clone().to(dev).requires_grad_(True); gw = torch.ones(256, device=dev, requires_grad=True)
    y = F.group_norm(a, 32, gw, None); (y * cot.to(dev)).sum().backward(); return y, a.grad, gw.grad
r = gn("cpu"); o = gn("cuda"); print("groupnorm y %.2e dx %.2e dw %.2e" % tuple(rel(a, c) for a, c in zip(o, r)))

import harness
from oracle.condgraph_oracle import build_oracle
from scan_b200.condgraph import build_condgraph
name = sys.argv[1] if len(sys.argv) > 1 else "c2f_small"
harness.CASES["_d"] = dict(harness.CASES[name], steps=["source"])
cfg = harness.build_case("_d")[0]
cpu = harness.run_case("_d", build_oracle(cfg), "oracle")
prod = harness.run_case("_d", build_condgraph(cfg, 256), "product", device="cuda")
print("flags", torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
for k in cpu:
    if any(t in k for t in ("dfeat", "grad/", "act_l", "feat_l", "loss")) and k in prod and cpu[k].shape == prod[k].shape:
        s = max(float(np.abs(cpu[k]).max()), 1e-30)
        print("%-50s %10.2e  (scale %.2e)" % (k, float(np.abs(prod[k].astype(np.float64) - cpu[k].astype(np.float64)).max()) / s, s))
