"""The CKA discriminator restatement (oracle.cka_discriminator_loss) against the golden vectors of the unmodified reference
module (tests/golden/cka.npz), and -- in the build container -- against the reference itself."""
import os

import numpy as np
import pytest
import torch

import cka_case
from oracle import condgraph_oracle as orc


def _oracle_run(name):
    from scan_b200.discriminator import FCOSDiscriminator_con
    d = cka_case.build(name)
    holder = FCOSDiscriminator_con(num_convs=d["num_convs"], num_classes=d["k"], grad_reverse_lambda=d["lam"], grl_applied_domain=d["grl_dom"])
    sd = cka_case.state_dict_for(holder, seed=5)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    feat = d["feat"].clone().requires_grad_(True)
    act = d["act"].clone().requires_grad_(True)
    loss = orc.cka_discriminator_loss(params, feat, act, d["target"], d["k"] - 1, d["num_convs"])
    loss.backward()
    rev_feat = d["grl_dom"] == "both" or d["domain"] == "target"
    rev_act = d["grl_dom"] == "both"
    res = {"loss": float(loss.detach()), "d_feature": (-d["lam"] * feat.grad if rev_feat else feat.grad).numpy(),
           "d_act": (-d["lam"] * act.grad if rev_act else act.grad).numpy()}
    for k, p in params.items():
        res["grad/" + k] = p.grad.numpy()
    return res


@pytest.mark.parametrize("name", list(cka_case.CASES))
def test_cka_oracle_matches_reference_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, "cka.npz"))
    got = _oracle_run(name)
    assert abs(got["loss"] - float(gold[name + "/loss"])) <= 1e-6 * max(1.0, abs(float(gold[name + "/loss"])))
    for k, v in got.items():
        if k == "loss":
            continue
        cka_case.check(v, gold[name + "/" + k + "#s"], gold[name + "/" + k + "#n"], 1e-5)


def test_state_dict_layout_matches_reference_names():
    from scan_b200.discriminator import FCOSDiscriminator_con
    m = FCOSDiscriminator_con(num_convs=4, num_classes=9)
    keys = set(m.state_dict())
    for i in range(4):
        assert {"dis_tower.%d.weight" % (3 * i), "dis_tower.%d.bias" % (3 * i), "dis_tower.%d.weight" % (3 * i + 1)} <= keys
    for c in range(8):
        assert {"classifier_cls_%d.0.weight" % c, "classifier_cls_%d.2.bias" % c} <= keys
    assert m.state_dict()["classifier_cls_0.0.weight"].shape == (128, 257, 3, 3)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 256, 4, 4), 1, act_maps=torch.zeros(1, 9, 4, 4))      # CPU tensors: no fallback


def test_block_structured_weights_reproduce_the_per_class_loop():
    """Host logic of the all-classes formulation (scan_b200/discriminator.py:_dense_weights), checked with plain torch on the CPU:
    ONE convolution of [features | all class maps] with the block-structured weight, then ONE with the block-diagonal weight,
    equals the reference's per-class Conv(cat(x, map_c)) -> ReLU -> Conv (fcos_head_discriminator_con.py:100-112)."""
    import torch.nn.functional as F
    from scan_b200.discriminator import FCOSDiscriminator_con
    torch.manual_seed(0)
    m = FCOSDiscriminator_con(num_convs=1, num_classes=4)
    m.load_state_dict(cka_case.state_dict_for(m, seed=9))
    x = torch.randn(2, 256, 6, 7)
    maps = torch.rand(2, 3, 6, 7)
    w1, b1, w2, b2 = m._dense_weights()
    assert w1.shape == (3 * 128, 256 + 3, 3, 3) and w2.shape == (3, 3 * 128, 3, 3)
    h = F.relu(F.conv2d(torch.cat([x, maps], 1), w1, b1, padding=1))
    logits = F.conv2d(h, w2, b2, padding=1)
    for c, block in enumerate(m.class_cond_map):
        want = block(torch.cat([x, maps[:, c:c + 1]], 1))
        assert torch.allclose(logits[:, c:c + 1], want, rtol=1e-4, atol=1e-5), c
