"""Error of the attention kernels against an fp64 reference, per implementation pair (forward x backward)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from scan_b200 import ops
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8728
dev = "cuda"
for regime, qs in (("peaked", 1.0), ("flat", 0.1)):
    g = torch.Generator(device=dev).manual_seed(1)
    q, k, v, cot = [torch.randn(m, 256, device=dev, generator=g) for _ in range(4)]
    q, k = q * qs, k * qs + 0.5
    v = v + 1.0
    qd, kd, vd = [t.double().requires_grad_(True) for t in (q, k, v)]
    att = torch.softmax(torch.bmm(qd.reshape(4, m, 64), kd.reshape(4, m, 64).transpose(1, 2)) * 0.25, dim=2)
    ref = torch.bmm(att, vd.reshape(4, m, 64)).reshape(m, 256)
    (ref * cot.double()).sum().backward()
    del att
    for fwd in ("t5", "ffma"):
        for bwd in ("t5", "ffma"):
            ops.ATTN_IMPL.update(fwd=fwd, bwd=bwd)
            a, b, c = [t.clone().requires_grad_(True) for t in (q, k, v)]
            out = ops.chunked_attention(a, b, c, 0.25, 0.0, 0)
            (out * cot).sum().backward()
            def e(x, y):
                return float((x.double() - y).abs().max() / y.abs().max())
            print(json.dumps({"regime": regime, "fwd": fwd, "bwd": bwd, "ctx": e(out, ref), "dq": e(a.grad, qd.grad), "dk": e(b.grad, kd.grad),
                              "dv": e(c.grad, vd.grad)}))
