"""Device timing of the "next" rows next to plain torch on the same GPU (development tool; torch side = the formulas of the oracle,
i.e. what the reference runs through cuDNN / ATen):
  f3  FCOSDiscriminator_con forward + backward on one FPN level (P3 of the 800x1344 pyramid, 8 images, 9 classes)
  f4  FCOSPostProcessor on 8 images, five levels, PRE_NMS_TOP_N = 1000
Usage: python tools/bench_next_rows.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scan_b200 import fcos_hooks  # noqa: E402
from scan_b200.discriminator import FCOSDiscriminator_con  # noqa: E402
from oracle import condgraph_oracle as orc  # noqa: E402  (development tool: the torch formulas to time against)


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def cka():
    torch.manual_seed(0)
    n, h, w, k = 8, 100, 168, 9
    m = FCOSDiscriminator_con(num_convs=4, num_classes=k, grad_reverse_lambda=0.02).cuda()
    feat = torch.randn(n, 256, h, w, device="cuda").requires_grad_(True)
    act = torch.softmax(torch.randn(n, k, h, w, device="cuda") * 2, 1).requires_grad_(True)
    state = {kk: v for kk, v in m.state_dict().items()}
    params = list(m.parameters())

    def mine():
        loss = m(feat, 0.9, act_maps=act, domain="source")
        torch.autograd.grad(loss, [feat, act] + params)
        return loss

    def ref():
        loss = orc.cka_discriminator_loss(dict(m.named_parameters()), feat, act, 0.9, k - 1, 4)
        torch.autograd.grad(loss, [feat, act] + params)
        return loss

    torch.backends.cudnn.benchmark = True
    l1, l2 = float(mine()), float(ref())
    print("CKA discriminator P3 x 8 images, 8 classes: loss %.6f (scan_b200) vs %.6f (torch)" % (l1, l2))
    print("  scan_b200 fwd+bwd %.2f ms   torch / cuDNN (tf32) fwd+bwd %.2f ms" % (timeit(mine), timeit(ref)))
    print("  peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 2 ** 30))
    del state


def postproc():
    rs = np.random.RandomState(5)
    shapes, strides, n, c = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)], [8, 16, 32, 64, 128], 8, 8
    probs, regs, ctrs = [], [], []
    for (h, w), s in zip(shapes, strides):
        p = rs.uniform(0.0, 0.04, (n, c, h, w))
        hot = rs.rand(n, c, h, w) < 0.02
        p[hot] = rs.uniform(0.06, 0.99, int(hot.sum()))
        probs.append(torch.from_numpy(p.astype(np.float32)).cuda())
        regs.append(torch.from_numpy(np.exp(rs.standard_normal((n, 4, h, w)) * 0.5 + np.log(s * 2.0)).astype(np.float32)).cuda())
        ctrs.append(torch.from_numpy(rs.standard_normal((n, 1, h, w)).astype(np.float32)).cuda())
    sizes = [(800, 1344)] * n
    pp = fcos_hooks.FCOSPostProcessor(0.05, 1000, 0.6, 100, 0, c + 1, mode="light", fpn_strides=strides)
    ms = timeit(lambda: pp(None, probs, regs, ctrs, sizes), n=20)
    t0 = time.perf_counter()
    orc.fcos_postprocess(shapes, strides, [p.cpu() for p in probs], [r.cpu() for r in regs], [t.cpu() for t in ctrs], sizes, 0.05, 1000, 0.6, 100)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print("FCOS post-processor, 8 images x 5 levels: scan_b200 %.3f ms (incl. the count read-back and BoxList construction); "
          "oracle restatement on the host %.0f ms" % (ms, cpu_ms))


if __name__ == "__main__":
    cka()
    postproc()
