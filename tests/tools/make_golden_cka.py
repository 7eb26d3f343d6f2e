"""Generate tests/golden/cka.npz from the UNMODIFIED reference FCOSDiscriminator_con (build container only).

    PYTHONPATH=. python tests/tools/make_golden_cka.py

For each case of tests/cka_case.py: the reference module with a seeded state dict runs on the CPU; stored are the loss, the
gradients w.r.t. the feature and the activation maps (i.e. after the gradient reversal) and w.r.t. every parameter, each as a
strided sample of <= 4096 elements plus its sum / sum of squares / max magnitude (cka_case.compact).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
import cka_case  # noqa: E402


def main():
    ref_shim.install()
    from fcos_core.modeling.discriminator.fcos_head_discriminator_con import FCOSDiscriminator_con
    out = {}
    for name in cka_case.CASES:
        d = cka_case.build(name)
        m = FCOSDiscriminator_con(num_convs=d["num_convs"], num_classes=d["k"], grad_reverse_lambda=d["lam"], grl_applied_domain=d["grl_dom"])
        m.load_state_dict(cka_case.state_dict_for(m, seed=5))
        feat = d["feat"].clone().requires_grad_(True)
        act = d["act"].clone().requires_grad_(True)
        loss = m(feat, d["target"], act_maps=act, domain=d["domain"])
        loss.backward()
        out[name + "/loss"] = np.array(float(loss.detach()), dtype=np.float64)
        tensors = {"d_feature": feat.grad.numpy(), "d_act": act.grad.numpy() if act.grad is not None else np.zeros(act.shape, dtype=np.float32)}
        for k, p in m.named_parameters():
            tensors["grad/" + k] = p.grad.numpy()
        for k, v in tensors.items():     # compact form (strided sample + norms): the full gradients would be 30 MB
            out[name + "/" + k + "#s"], out[name + "/" + k + "#n"] = cka_case.compact(v)
        print(name, float(loss.detach()), float(feat.grad.abs().max()), float(np.abs(tensors["d_act"]).max()))
    meta = "reference=/root/reference FCOSDiscriminator_con torch=%s" % torch.__version__
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cka.npz"), __meta__=np.array(meta), **out)


if __name__ == "__main__":
    main()
