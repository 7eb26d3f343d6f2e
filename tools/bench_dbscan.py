import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from scan_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 36000
far = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
pts = torch.randn(npts, 256, device="cuda", generator=g) * 0.15
pts[: npts // 20] *= far
for _ in range(2):
    lab, info = ops.dbscan_points(pts, 3.0)
torch.cuda.synchronize()
print(info.tolist())
