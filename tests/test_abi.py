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


def test_rows_layout_views_and_join_host_logic():
    """The zero-copy plumbing of the rows layout is plain torch view arithmetic: a level's slice of the rows matrix IS a
    channels-last [N,C,H,W] tensor, and join_rows recognises adjacent views of one buffer (no copy) vs anything else."""
    import torch
    from scan_b200 import ops
    shapes = [(5, 7), (3, 4), (2, 2)]
    geo = ops.Geometry(shapes, [8, 16, 32], 2)
    rows = torch.arange(geo.R * ops.C, dtype=torch.float32).reshape(geo.R, ops.C)
    views = ops.level_views(geo, rows)
    for l, (h, w) in enumerate(shapes):
        v = views[l]
        assert v.shape == (2, ops.C, h, w) and v.is_contiguous(memory_format=torch.channels_last)
        assert v.data_ptr() == rows.data_ptr() + geo.row_off[l] * ops.C * 4
        # the reference's flattening: features[l].permute(0, 2, 3, 1).reshape(-1, C)   (loss.py:440)
        assert torch.equal(v.permute(0, 2, 3, 1).reshape(-1, ops.C), rows[geo.row_off[l]:geo.row_off[l + 1]])
        assert ops.nhwc_dense(v) is v
    joined = ops._JoinRows.apply(geo, *views)
    assert joined.data_ptr() == rows.data_ptr() and torch.equal(joined, rows)            # adjacent views: zero-copy
    separate = [v.clone(memory_format=torch.contiguous_format) for v in views]            # NCHW copies: concatenating path
    joined2 = ops._JoinRows.apply(geo, *separate)
    assert joined2.data_ptr() != separate[0].data_ptr() and torch.equal(joined2, rows)
    # gradients come back as per-level views of the incoming [R, C] gradient
    leaves = [s.clone().requires_grad_(True) for s in separate]
    out = ops._JoinRows.apply(geo, *leaves)
    cot = torch.randn(geo.R, ops.C)
    (out * cot).sum().backward()
    for l, (h, w) in enumerate(shapes):
        want = cot[geo.row_off[l]:geo.row_off[l + 1]].view(2, h, w, ops.C).permute(0, 3, 1, 2)
        assert torch.equal(leaves[l].grad, want)


def test_launch_table_covers_every_compute_entry_point():
    """bench.py's gpu_launches is counted from _lib.LAUNCHES: every entry point that launches kernels must be listed."""
    no_kernels = {"scan_abi_version", "scan_strerror", "scan_last_cuda_error", "scan_init", "scan_condconv_num_partials"}
    for name in _lib.SIGNATURES:
        if name in no_kernels or name.endswith("_bytes") or name.endswith("_floats") or name.endswith("_num_partials"):
            continue
        assert name in _lib.LAUNCHES and _lib.LAUNCHES[name] >= 1, name
