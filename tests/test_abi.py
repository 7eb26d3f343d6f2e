"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/scan_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from scan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "scan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scan_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), "libscan_b200.so does not export %s" % name
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)


def test_abi_version_and_error_strings():
    L = _lib.lib()
    assert L.scan_abi_version() == 1
    assert L.scan_strerror(0) == b"ok"
    assert b"invalid" in L.scan_strerror(-1)
    assert L.scan_sample_workspace_bytes(1000) > 4000
    assert L.scan_dbscan_workspace_bytes(1024) > 1024 * 1024


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.ScanLevels) == 4 * (2 + 3 * 8)
    assert ctypes.sizeof(_lib.ScanSampleMeta) == 4 * (4 + 5 * 8)


def test_product_refuses_cpu_tensors():
    import torch
    from scan_b200 import ops
    with pytest.raises(RuntimeError):
        ops._ptr(torch.zeros(4))
    from scan_b200.condgraph import build_condgraph
    from scan_b200.config import scan_cfg
    m = build_condgraph(scan_cfg("c2f"), 256)
    feats = [torch.zeros(1, 256, 4, 4) for _ in range(5)]
    with pytest.raises(RuntimeError):
        m.eval()(None, feats)


def test_state_dict_matches_reference_contract():
    """keys and shapes of the shipped C2F config (3 636 233 parameters)."""
    from scan_b200.condgraph import build_condgraph
    from scan_b200.config import scan_cfg
    m = build_condgraph(scan_cfg("c2f"), 256)
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 3636233
    assert tuple(sd["prototype"].shape) == (9, 256, 3)
    assert tuple(sd["head_out.middle_tower.0.weight"].shape) == (256, 265, 3, 3)
    assert tuple(sd["cond_nx1.weight"].shape) == (256, 512, 3, 1)
    assert tuple(sd["cond_rnn.weight_hh_l1"].shape) == (512, 512)
    for k in ("multihead_attn.linear_k.weight", "multihead_attn.layer_norm.bias", "proto_cls.weight", "cond_2.weight",
              "head_in.middle_tower.4.weight"):
        assert k in sd
    from oracle.condgraph_oracle import build_oracle
    assert list(build_oracle(scan_cfg("c2f")).state_dict().keys()) == list(sd.keys())
