"""f4: the oracle restatement of FCOSPostProcessor against the golden vectors generated from the unmodified reference class
(tests/tools/make_golden_postproc.py; its NMS stub is the oracle's nms_indices, see that script)."""
import os

import numpy as np
import pytest

import postproc_case
from oracle import condgraph_oracle as orc


@pytest.mark.parametrize("name", list(postproc_case.CASES))
def test_oracle_postprocessor_matches_reference_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, "postproc.npz"))
    d = postproc_case.build(name)
    res = orc.fcos_postprocess(d["shapes"], d["strides"], d["probs"], d["regs"], d["ctrs"], d["sizes"], d["thr"], d["top_n"],
                               d["nms_thr"], d["post_n"])
    for i, (b, s, l) in enumerate(res):
        want = gold["%s/img%d" % (name, i)]
        got = postproc_case.canonical(b.numpy(), s.numpy(), l.numpy())
        assert got.shape == want.shape, (got.shape, want.shape)
        assert np.array_equal(got[:, 0], want[:, 0])
        assert np.abs(got[:, 1:] - want[:, 1:]).max() <= 1e-5 * max(1.0, np.abs(want[:, 1:]).max())
        if name == "c8":     # no level exceeded pre_nms_top_n: the reference's own output order is defined and must match
            assert np.array_equal(l.numpy(), gold["%s/img%d_labels_in_order" % (name, i)])
            assert np.allclose(s.numpy(), gold["%s/img%d_scores_in_order" % (name, i)], rtol=1e-6)
