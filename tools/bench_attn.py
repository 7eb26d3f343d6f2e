import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from scan_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8728
q, k, v = [torch.randn(m, 256, device="cuda", generator=g).requires_grad_(True) for _ in range(3)]
cot = torch.randn(m, 256, device="cuda", generator=g)
def t(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    fwd = t(lambda: ops.chunked_attention(q, k, v, 0.25, 0.1, 5))
def fb():
    out = ops.chunked_attention(q, k, v, 0.25, 0.1, 5); out.backward(cot); q.grad = k.grad = v.grad = None
tot = t(fb)
flops = 4 * 2 * 2 * m * m * 64
print(json.dumps({"impl": os.environ.get("SCAN_B200_ATTN_TC", "0"), "M": m, "fwd_ms": fwd, "fwd_bwd_ms": tot, "fwd_TFLOPs": flops / fwd / 1e9}))
