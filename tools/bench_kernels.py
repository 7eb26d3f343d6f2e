"""Micro-benchmarks of individual C-ABI entry points (CUDA events, L2 flushed between iterations)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from scan_b200 import ops

FULL = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
dev = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
geo = ops.Geometry(FULL, [8, 16, 32, 64, 128], n)
g = torch.Generator(device=dev).manual_seed(0)
rows = torch.relu(torch.randn(geo.R, 256, device=dev, generator=g))
w = torch.randn(9, 256, device=dev, generator=g) * 0.05
labels = torch.randint(0, 9, (geo.R,), device=dev, generator=g)
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)

def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

res = {}
fwd_bytes = geo.R * (1024 + 36 + 8)
for impl in (0, 1):
    ops.CONDCONV_IMPL["impl"] = impl
    try:
        ms = timeit(lambda: ops.condconv(geo, rows, w, None, 9, 0, labels, 1.0))
        res["condconv_fwd_impl%d" % impl] = {"ms": ms, "GBps": fwd_bytes / ms / 1e6, "frac_of_6457": fwd_bytes / ms / 1e6 / 6457.1}
    except Exception as e:
        res["condconv_fwd_impl%d" % impl] = repr(e)[:200]
ops.CONDCONV_IMPL["impl"] = 0
if os.environ.get("SCAN_B200_CC_KMAJOR") == "1":
    rk = rows.view(geo.R, 8, 32).permute(1, 0, 2).contiguous().view(geo.R, 256)
    a_ref, _, _ = ops.condconv(geo, rk, w, None, 9, 0, labels, 1.0)
    ms = timeit(lambda: ops.condconv(geo, rk, w, None, 9, 0, labels, 1.0))
    res["condconv_fwd_kmajor"] = {"ms": ms, "GBps": fwd_bytes / ms / 1e6, "frac_of_6457": fwd_bytes / ms / 1e6 / 6457.1}
    print(json.dumps(res, indent=1)); sys.exit(0)
rr = rows.clone().requires_grad_(True); ww = w.clone().requires_grad_(True)
acts, loss, _ = ops.condconv(geo, rr, ww, None, 9, 0, labels, 1.0)
cots = [torch.randn_like(a) for a in acts]
def bwd():
    torch.autograd.backward([loss] + acts, [torch.ones_like(loss)] + cots, retain_graph=True)
    rr.grad = None; ww.grad = None
ms = timeit(bwd)
bwd_bytes = geo.R * (2048 + 72 + 8)
res["condconv_bwd"] = {"ms": ms, "GBps": bwd_bytes / ms / 1e6, "frac_of_6457": bwd_bytes / ms / 1e6 / 6457.1}
feats = [torch.randn(n, 256, h, w_, device=dev, generator=g) for h, w_ in FULL]
ms = timeit(lambda: ops.pack_rows(geo, feats))
res["pack_rows"] = {"ms": ms, "GBps": 2 * geo.R * 1024 / ms / 1e6, "frac_of_6457": 2 * geo.R * 1024 / ms / 1e6 / 6457.1}
m = 8728 * n // 8
q, k, v = [torch.randn(m, 256, device=dev, generator=g) for _ in range(3)]
ms = timeit(lambda: ops.chunked_attention(q, k, v, 0.25), 5)
res["attn_fwd_M%d" % m] = {"ms": ms, "TFLOPs": 4 * 2 * 2 * m * m * 64 / ms / 1e9}
for npts in (8192, 36000):
    pts = torch.randn(npts, 256, device=dev, generator=g) * 0.15
    pts[: npts // 20] *= 30
    ms = timeit(lambda: ops.dbscan_points(pts, 3.0), 3)
    lab, info = ops.dbscan_points(pts, 3.0)
    res["dbscan_points_n%d" % npts] = {"ms": ms, "info": info.tolist()}
print(json.dumps(res, indent=1))
