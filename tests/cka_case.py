"""Seeded inputs of the CKA discriminator parity cases (shared by tests/tools/make_golden_cka.py, the oracle test and the GPU
test): one FPN level's features, softmax-like activation maps, a seeded state dict."""
import numpy as np
import torch

# name -> (images, H, W, FCOS NUM_CLASSES (fg + bg), num_convs, grad_reverse_lambda, grl domain, domain, target label)
CASES = {
    "c9_source": (2, 13, 21, 9, 2, 0.02, "both", "source", 0.9),
    "c9_target": (2, 9, 14, 9, 2, 0.02, "both", "target", 0.1),
    "k2_target_only": (3, 7, 11, 2, 1, -1.0, "target", "target", 1),
}


def build(name):
    n, h, w, k, num_convs, lam, grl_dom, domain, target = CASES[name]
    rs = np.random.RandomState({"c9_source": 11, "c9_target": 12, "k2_target_only": 13}[name])
    feat = torch.from_numpy(rs.standard_normal((n, 256, h, w)).astype(np.float32))
    logits = rs.standard_normal((n, k, h, w)).astype(np.float32) * 2.0
    e = np.exp(logits - logits.max(1, keepdims=True))
    act = torch.from_numpy((e / e.sum(1, keepdims=True)).astype(np.float32))
    return dict(n=n, h=h, w=w, k=k, num_convs=num_convs, lam=lam, grl_dom=grl_dom, domain=domain, target=target, feat=feat, act=act)


def state_dict_for(module, seed):
    """Seeded parameters large enough that every path matters (the reference initialises with std 0.01 and zero biases)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in module.state_dict().items():
        if k.endswith("weight") and v.dim() == 4:
            sd[k] = torch.randn(v.shape, generator=g) * (2.0 / (v.shape[1] * 9) ** 0.5)
        elif k.endswith("weight"):
            sd[k] = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
    return sd


SAMPLE = 4096


def compact(arr):
    """Small fixture form of a tensor: a strided sample of at most SAMPLE elements plus (sum, sum of squares, max |.|) in fp64."""
    a = np.asarray(arr, dtype=np.float32).reshape(-1)
    stride = max(1, (a.size + SAMPLE - 1) // SAMPLE)
    a64 = a.astype(np.float64)
    return a[::stride].copy(), np.array([a64.sum(), (a64 * a64).sum(), np.abs(a64).max() if a.size else 0.0])


def check(arr, sample, norms, rtol):
    """Largest error of `arr` against a compact fixture, relative to the reference tensor's max magnitude."""
    got_s, got_n = compact(arr)
    scale = max(float(norms[2]), 1e-12)
    assert got_s.shape == sample.shape
    err = float(np.abs(got_s.astype(np.float64) - sample.astype(np.float64)).max()) / scale
    # the norms catch what the sample misses: relative difference of the L2 norms, and of the sums relative to the L1 scale
    err = max(err, abs(np.sqrt(got_n[1]) - np.sqrt(norms[1])) / max(np.sqrt(norms[1]), 1e-12))
    assert err <= rtol, err
    return err
