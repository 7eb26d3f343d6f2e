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
