"""Device timing of the tower convolution kernel (csrc/tower.cu) next to cuDNN (NHWC, tf32) on the bench geometry.
Usage: python tools/bench_conv.py [n_images]"""
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scan_b200 import ops  # noqa: E402

FULL = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
STRIDES = [8, 16, 32, 64, 128]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.backends.cudnn.benchmark = True
    geo = ops.Geometry(FULL, STRIDES, n)
    x = torch.randn(geo.R, 256, device="cuda")
    w = torch.randn(256, 256, 3, 3, device="cuda") * 0.02
    flops = 2.0 * geo.R * 256 * 256 * 9
    hi, lo = ops.conv3x3_pack(w, False, True)
    xlo = ops.tf32_residual(x)
    out = torch.empty(geo.R, 256, device="cuda")
    for cg in (1, 2):
        ms = timeit(lambda: ops.conv3x3_rows_raw(geo, x, hi, 256, out=out, cta_group=cg))
        print("tower conv tf32   cta_group %d: %.3f ms  %.0f TFLOP/s" % (cg, ms, flops / ms / 1e9))
    ms = timeit(lambda: ops.conv3x3_rows_raw(geo, x, hi, 256, out=out, x_lo=xlo, packed_lo=lo, cta_group=2), 5)
    print("tower conv 3xTF32 cta_group 2: %.3f ms  %.0f TFLOP/s (algorithmic)" % (ms, flops / ms / 1e9))
    ms = timeit(lambda: ops.conv3x3_pack(w, False, False))
    print("weight pack: %.4f ms" % ms)
    dy = torch.randn(geo.R, 256, device="cuda")
    gw = torch.empty(256, 256, 3, 3, device="cuda")
    ms = timeit(lambda: ops.conv3x3_wgrad_raw(geo, x, dy, out=gw))
    print("tower wgrad tf32 (+reduce): %.3f ms  %.0f TFLOP/s" % (ms, flops / ms / 1e9))
    xs = ops.level_views(geo, x)
    wc = w.contiguous(memory_format=torch.channels_last)
    ms = timeit(lambda: [torch.nn.functional.conv2d(t, wc, None, padding=1) for t in xs])
    print("cuDNN fprop (5 launches, tf32 %s): %.3f ms  %.0f TFLOP/s" % (torch.backends.cudnn.allow_tf32, ms, flops / ms / 1e9))
    xs_g = [t.detach().requires_grad_(True) for t in xs]
    wg = wc.detach().requires_grad_(True)
    ys = [torch.nn.functional.conv2d(t, wg, None, padding=1) for t in xs_g]
    gs = [torch.randn_like(y) for y in ys]
    ms = timeit(lambda: torch.autograd.grad(ys, xs_g, gs, retain_graph=True))
    print("cuDNN dgrad: %.3f ms  %.0f TFLOP/s" % (ms, flops / ms / 1e9))
    ms = timeit(lambda: torch.autograd.grad(ys, [wg], gs, retain_graph=True))
    print("cuDNN wgrad: %.3f ms  %.0f TFLOP/s" % (ms, flops / ms / 1e9))


if __name__ == "__main__":
    main()
