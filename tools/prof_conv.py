"""Two launches of each tower-convolution kernel on the bench geometry (for ncu)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scan_b200 import ops  # noqa: E402

FULL = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
geo = ops.Geometry(FULL, [8, 16, 32, 64, 128], int(sys.argv[1]) if len(sys.argv) > 1 else 16)
x = torch.randn(geo.R, 256, device="cuda")
dy = torch.randn(geo.R, 256, device="cuda")
w = torch.randn(256, 256, 3, 3, device="cuda") * 0.02
hi, _ = ops.conv3x3_pack(w, False, False)
for _ in range(2):
    ops.conv3x3_rows_raw(geo, x, hi, 256, cta_group=2)
    ops.conv3x3_wgrad_raw(geo, x, dy)
torch.cuda.synchronize()
