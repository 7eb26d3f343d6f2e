"""Live pin of the oracle against the unmodified reference (build container only: needs
/root/reference).  On the GPU box these are skipped and the committed golden fixtures take over."""
import numpy as np
import pytest

import harness

pytestmark = []


@pytest.mark.reference
@pytest.mark.parametrize("name", ["c2f_small", "gcn_small", "sim10k_small"])
def test_reference_still_matches_golden(name, golden_dir):
    import os
    from oracle import ref_shim
    from scan_b200.config import to_plain
    _, _, BoxList = ref_shim.reference_modules()
    cfg = harness.build_case(name)[0]
    ref = ref_shim.build_reference(to_plain(cfg))
    got = harness.run_case(name, ref, "reference", boxlist_cls=BoxList)
    want = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    want.pop("__meta__")
    bad = harness.compare(got, want, rtol=1e-5)
    assert not bad, "\n".join(bad[:20])


def test_floor_linspace_matches_numpy():
    from oracle.condgraph_oracle import floor_linspace
    rs = np.random.RandomState(0)
    for _ in range(3000):
        n_neg = int(rs.randint(1, 200000))
        n_pos = int(rs.randint(0, 5000))
        want = np.floor(np.linspace(0, n_neg - 2, n_pos)).astype(int)
        assert np.array_equal(floor_linspace(n_neg - 2, n_pos), want)


def test_dbscan_restatement_matches_sklearn():
    from sklearn.cluster import DBSCAN
    from oracle.condgraph_oracle import dbscan_labels
    rs = np.random.RandomState(1)
    for trial in range(12):
        n = int(rs.randint(50, 600))
        if trial % 2:
            # planar cloud embedded in 256-d: many border points
            p2 = rs.uniform(0, 30, (n, 2))
            basis = np.linalg.qr(rs.standard_normal((256, 2)))[0]
            x = (p2 @ basis.T).astype(np.float32)
        else:
            centers = rs.standard_normal((4, 256)) * 3
            x = (centers[rs.randint(0, 4, n)] + rs.standard_normal((n, 256)) * 0.17).astype(np.float32)
        want = DBSCAN(eps=3.0).fit_predict(x)
        assert np.array_equal(dbscan_labels(x, 3.0), want)


def test_dbscan_c_oracle_matches_sklearn():
    from sklearn.cluster import DBSCAN
    from oracle.condgraph_oracle import dbscan_labels_c, _c_dbscan
    if not _c_dbscan():
        import subprocess, os
        subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")])
        import oracle.condgraph_oracle as o
        o._C_DBSCAN = None
    rs = np.random.RandomState(7)
    for trial in range(10):
        n = int(rs.randint(1, 900))
        if trial % 2:
            p2 = rs.uniform(0, 40, (n, 2))
            basis = np.linalg.qr(rs.standard_normal((256, 2)))[0]
            x = (p2 @ basis.T).astype(np.float32)
        else:
            centers = rs.standard_normal((4, 256)) * 3
            x = (centers[rs.randint(0, 4, n)] + rs.standard_normal((n, 256)) * 0.17).astype(np.float32)
        assert np.array_equal(dbscan_labels_c(x, 3.0), DBSCAN(eps=3.0).fit_predict(x))
