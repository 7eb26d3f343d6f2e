"""Shared parity harness: drives ANY implementation of the middle head (the unmodified reference
through oracle/ref_shim.py, the CPU oracle, or the CUDA product) through the same seeded sequence of
steps and reduces what it produced to a flat dict of numpy arrays.

`tests/tools/make_golden.py` runs it on the reference and commits the result under tests/golden/;
the tests run it on the oracle (CPU) and on scan_b200 (GPU) and compare.
"""
import hashlib

import numpy as np
import torch

from scan_b200.config import scan_cfg
from scan_b200.fixtures import fixture_state_dict
from scan_b200.synthetic import make_workload
from scan_b200.structures import BoxList

SMALL_SHAPES = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]
SMALL_HW = (200, 336)

# name -> dict(cfg=(preset, overrides), n=images, steps=[...], fixture=(kernel_gain, gn_gain))
CASES = {
    "c2f_small": dict(cfg=("c2f", {}), n=2, steps=["source"] * 5 + ["target", "eval"], fixture=(1.0, 1.0)),
    "c2f_eps45": dict(cfg=("c2f", {"DBSCAN_EPS": 4.5}), n=2, steps=["source", "target"], fixture=(2.0, 1.0)),
    "c2f_selftrain": dict(cfg=("c2f", {"GCN_SELF_TRAINING": True, "TRANSFER_CFG": ("NODES", "ADJ", "PROTOTYPE")}),
                          n=2, steps=["source", "target"], fixture=(1.0, 1.0)),
    "sim10k_small": dict(cfg=("sim10k", {}), n=2, steps=["source", "source", "target", "eval"], fixture=(4.0, 1.0)),
    "gcn_small": dict(cfg=("c2f", {"GLOBAL_GCN": False}), n=2, steps=["source", "source", "target", "eval"],
                      fixture=(1.0, 1.0)),
    "gcn_no_norm": dict(cfg=("c2f", {"GLOBAL_GCN": False, "GCN_EDGE_NORM": "NO", "GCN_OUT_ACTIVATION": "tanh",
                                     "GCN_SHORTCUT": True}), n=1, steps=["source", "eval"], fixture=(1.0, 1.0)),
    "nornn_p3": dict(cfg=("c2f", {"USE_RNN": None}), n=1, steps=["source"] * 4 + ["eval"], fixture=(0.05, 1.0)),
    "p1_bias": dict(cfg=("c2f", {"USE_RNN": None, "PROTO_ITER": 1, "COND_WITH_BIAS": True, "COSINE_UPDATE_ON": False}),
                    n=1, steps=["source", "source", "eval"], fixture=(30.0, 1.0)),
    "score_thr": dict(cfg=("c2f", {"TARGET_SAMPLING_CFG": "score_threshold"}), n=2, steps=["source", "target"],
                      fixture=(3.0, 1.0), plabel_th=0.97),
}


def build_case(name, boxlist_cls=BoxList):
    case = CASES[name]
    preset, over = case["cfg"]
    cfg = scan_cfg(preset, **over)
    if "plabel_th" in case:
        cfg.SOLVER.MIDDLE_HEAD.PLABEL_TH = (case["plabel_th"],)
    num_fg = cfg.MODEL.FCOS.NUM_CLASSES - 1
    n = case["n"]
    shapes, hw, nb = (None, (800, 1344), 18) if case.get("full") else (SMALL_SHAPES, SMALL_HW, 6)
    src_feats, src_targets = make_workload(n, num_fg, seed=1234, boxes_per_image=nb, level_shapes=shapes,
                                           image_hw=hw, boxlist_cls=boxlist_cls)
    tgt_feats, _ = make_workload(n, num_fg, seed=4321, boxes_per_image=nb, level_shapes=shapes,
                                 image_hw=hw, boxlist_cls=boxlist_cls)
    return cfg, case, src_feats, src_targets, tgt_feats


def cotangents(shapes, seed):
    rs = np.random.RandomState(seed)
    out = []
    for s in shapes:
        g = rs.standard_normal(tuple(s)).astype(np.float32) / float(np.prod(s)) * 16.0
        out.append(torch.from_numpy(g))
    return out


def _sample(t, limit=1024):
    a = t.detach().float().cpu().reshape(-1).numpy()
    if a.size <= limit:
        return a.copy()
    stride = a.size // limit
    return a[::stride][:limit].copy()


def _row_lookup(feats_in):
    table = {}
    for l, f in enumerate(feats_in):
        rows = f.detach().cpu().permute(0, 2, 3, 1).reshape(-1, f.shape[1]).contiguous().numpy()
        for r in range(rows.shape[0]):
            table[hashlib.md5(rows[r].tobytes()).digest()] = (l, r)
    return table


def _rows_of(points, table):
    pts = points.detach().cpu().contiguous().numpy()
    lv = np.zeros(len(pts), np.int64)
    rw = np.zeros(len(pts), np.int64)
    for i in range(len(pts)):
        lv[i], rw[i] = table[hashlib.md5(pts[i].tobytes()).digest()]
    return lv, rw


class ReferenceProbe(object):
    """Wraps the reference's PrototypeComputation instance to expose what it returns."""

    def __init__(self, module):
        self.m = module
        self.inner = module.prototype_evaluator
        self.rec = {}
        module.prototype_evaluator = self
        inner = self.inner
        orig_db = inner.DBSCAN_batch_cpu
        probe = self

        def db(act, feat):
            y = orig_db(act, feat)
            probe.rec.setdefault("dbscan_masks", []).append(y.clone())
            return y

        inner.DBSCAN_batch_cpu = db

    def __call__(self, locations=None, features=None, targets=None):
        for f in features:
            if f.requires_grad:
                f.retain_grad()       # d(loss)/d(head_in output): what the hot path's backward produces
        out = self.inner(locations, features, targets)
        self.rec["features_in"] = list(features)  # the reference later overwrites the list entries in place (condgraph.py:382)
        # cloned: with GLOBAL_GCN=False the reference overwrites the sampled rows in place (condgraph.py:413)
        self.rec["points"] = out[0].detach().clone() if out[0] is not None else None
        self.rec["point_labels"], self.rec["labels"] = out[1], out[2]
        return out

    def __getattr__(self, k):
        return getattr(self.inner, k)


def run_case(name, module, impl, device="cpu", boxlist_cls=BoxList, load_fixture=True):
    """impl: 'reference' | 'oracle' | 'product'.  Returns dict[str, np.ndarray]."""
    cfg, case, src_feats, src_targets, tgt_feats = build_case(name, boxlist_cls)
    if load_fixture:
        kg, gg = case["fixture"]
        sd = fixture_state_dict(module, seed=99, kernel_gain=kg, gn_gain=gg)
        module.load_state_dict(sd)
    module.to(device) if impl != "reference" else None
    if device != "cpu":
        # parity runs compare against an fp32 CPU reference: keep torch's own conv / matmul kernels (the towers,
        # which stay on cuDNN/cuBLAS) in true fp32; the only tf32 arithmetic left is the tcgen05 conditional conv
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
    # dropout off everywhere: parity runs are deterministic
    if hasattr(module, "multihead_attn"):
        a = module.multihead_attn
        if impl == "reference":
            a.dropout.p = 0.0
            a.dot_product_attention.dropout.p = 0.0
        else:
            a.p_drop = 0.0
    probe = ReferenceProbe(module) if impl == "reference" else None
    res = {}
    params = dict(module.named_parameters())
    for i, step in enumerate(case["steps"]):
        pre = "s%d/" % i
        feats_cpu = src_feats if step == "source" else tgt_feats
        feats = [f.clone().to(device).requires_grad_(step != "eval") for f in feats_cpu]
        for p in params.values():
            p.grad = None
        if probe:
            probe.rec = {}
        if step == "eval":
            module.eval()
            with torch.no_grad():
                out = module(None, feats)
        else:
            module.train()
            if step == "source":
                tg = [t.to(device) for t in src_targets] if device != "cpu" else src_targets
                out = module(None, feats, targets=tg, mode="source")
            else:
                out = module(None, feats, targets=None, mode="target", forward_target=True)
        out_feats, loss_graph, act_loss, acts = out
        fin = (probe.rec.get("features_in") if probe else module.last.get("features_in")) if step != "eval" else None
        if fin is not None and not probe:
            for f in fin:
                if f.requires_grad:
                    f.retain_grad()
        # ---- intermediate results ----
        if impl == "reference":
            rec = probe.rec
            if step == "source":
                for l, lab in enumerate(rec["labels"]):
                    res[pre + "labels_l%d" % l] = lab.cpu().numpy().astype(np.int64)
            if step != "eval" and rec.get("points") is not None:
                table = _row_lookup(rec["features_in"])
                lv, rw = _rows_of(rec["points"], table)
                res[pre + "node_level"], res[pre + "node_rows"] = lv, rw
                res[pre + "node_labels"] = rec["point_labels"].cpu().numpy().astype(np.int64)
            for l, mk in enumerate(rec.get("dbscan_masks", [])):
                res[pre + "dbscan_mask_l%d" % l] = mk.cpu().numpy().astype(np.uint8)
        else:
            last = module.last
            if step == "source":
                for l, lab in enumerate(last["labels"]):
                    res[pre + "labels_l%d" % l] = lab.detach().cpu().numpy().astype(np.int64)
            if step != "eval" and last.get("node_rows") is not None:
                res[pre + "node_level"] = last["node_level"].detach().cpu().numpy().astype(np.int64)
                res[pre + "node_rows"] = last["node_rows"].detach().cpu().numpy().astype(np.int64)
                res[pre + "node_labels"] = last["node_labels"].detach().cpu().numpy().astype(np.int64)
            for l, mk in enumerate(last.get("dbscan_masks", [])):
                res[pre + "dbscan_mask_l%d" % l] = mk.detach().cpu().numpy().astype(np.uint8)
        # ---- losses ----
        node_loss = transfer = None
        if loss_graph is not None:
            node_loss, transfer = loss_graph
        res[pre + "has_loss_graph"] = np.array(int(loss_graph is not None))
        for key, v in (("node_loss", node_loss), ("transfer_loss", transfer), ("act_loss", act_loss)):
            if torch.is_tensor(v):
                res[pre + key] = np.array(float(v.detach().cpu()), dtype=np.float64)
        # ---- maps / features ----
        for l in range(len(acts)):
            res[pre + "act_l%d" % l] = _sample(acts[l])
            res[pre + "feat_l%d" % l] = _sample(out_feats[l])
        # ---- backward ----
        if step != "eval":
            gf = cotangents([f.shape for f in out_feats], 1000 + i)
            ga = cotangents([a.shape for a in acts], 2000 + i)
            direct = sum((a * g.to(device)).sum() for a, g in zip(acts, ga))
            for v in (node_loss, transfer, act_loss):
                if torch.is_tensor(v):
                    direct = direct + v
            # "dfin_direct": d(losses + act-map cotangent)/d(features_in) -- the part of the backward that runs entirely
            # through the hot path (conditional conv, gather, attention ...) and through no tower ReLU; also exercises
            # the double backward call of engine/trainer.py:299,343 (retain_graph=True)
            if fin is not None and any(f.requires_grad for f in fin):
                gd = torch.autograd.grad(direct, fin, retain_graph=True, allow_unused=True)
                for l, g_ in enumerate(gd):
                    if g_ is not None:
                        res[pre + "dfin_direct_l%d" % l] = _sample(g_)
            total = direct + sum((f * g.to(device)).sum() for f, g in zip(out_feats, gf))
            total.backward()
            for l, f in enumerate(feats):
                res[pre + "dfeat_l%d" % l] = _sample(f.grad)
            if fin is not None:
                for l, f in enumerate(fin):
                    if f.grad is not None:
                        res[pre + "dfin_l%d" % l] = _sample(f.grad)
            for pname, p in params.items():
                if p.grad is not None:
                    res[pre + "grad/" + pname] = _sample(p.grad)
                    res[pre + "gradnorm/" + pname] = np.array(float(p.grad.detach().double().norm().cpu()))
        res[pre + "prototype"] = module.prototype.detach().cpu().numpy().copy()
    if probe:
        module.prototype_evaluator = probe.inner
    return res


INT_KEYS = ("labels_l", "node_level", "node_rows", "node_labels", "dbscan_mask_l", "has_loss_graph")


# d(loss)/d(bias of the LAST manifestation layer) is analytically ZERO (cond_nx1.bias with USE_RNN, cond_2.bias otherwise) when it
# feed a softmax without COND_WITH_BIAS: the bias shifts every class logit of a pixel by the same amount.
# What the reference returns there is pure cancellation noise (|g| ~ 1e-7), compared with an absolute bound.
NOISE_KEYS = ("grad/cond_nx1.bias", "gradnorm/cond_nx1.bias", "grad/cond_2.bias", "gradnorm/cond_2.bias")
NOISE_ATOL = 5e-6


def compare(got, want, rtol=1e-3, atol_scale=1e-3, skip=(), device_run=False):
    """Bit-exact for integer results, rtol (relative to the tensor's max magnitude) for floats.
    Returns list of human-readable mismatches."""
    bad = []
    for k, w in want.items():
        if any(s in k for s in skip):
            continue
        if k not in got:
            bad.append("missing %s" % k)
            continue
        g = got[k]
        if any(t in k for t in INT_KEYS):
            if g.shape != w.shape or not np.array_equal(g, w):
                bad.append("int mismatch %s (%s vs %s)" % (k, g.shape, w.shape))
            continue
        if g.shape != w.shape:
            bad.append("shape mismatch %s %s vs %s" % (k, g.shape, w.shape))
            continue
        scale = max(float(np.abs(w).max()) if w.size else 0.0, 1e-30)
        err = float(np.abs(g.astype(np.float64) - w.astype(np.float64)).max()) if w.size else 0.0
        if any(t in k for t in NOISE_KEYS) and scale < NOISE_ATOL:
            if err > NOISE_ATOL:
                bad.append("noise-floor mismatch %s: max|d|=%.3e" % (k, err))
            continue
        if err > rtol * scale + atol_scale * 1e-6:
            # Gradients that pass through a ReLU (head_in / head_out towers, both torch ops) are discontinuous: an
            # activation within ~1e-6 of zero gets the opposite mask on the GPU and on the CPU, which changes a FEW
            # gradient entries by O(1) relative amounts (measured: tests/tools/diag_grad.py; the same happens between the
            # reference on CPU and the reference on GPU).  Such tensors pass if the outliers are sparse and the
            # relative L2 error is small; everything else must meet the max-norm bound.
            d = np.abs(g.astype(np.float64) - w.astype(np.float64))
            frac = float((d > rtol * scale).mean())
            rel_l2 = float(np.linalg.norm(d) / max(np.linalg.norm(w.astype(np.float64)), 1e-30))
            relu_path = ("dfeat_l" in k) or ("dfin_l" in k) or ("dfin_direct" in k) or ("head_in." in k) or ("head_out." in k) or ("proto_cls_hidden" in k)
            if relu_path and frac <= 0.02 and rel_l2 <= 3 * rtol:
                continue
            # d(input features) and the head_in parameter gradients are produced by torch's OWN backward (cuDNN
            # backward-data / backward-filter, GroupNorm) from d(features_in) = "dfin", which is what the scan_b200
            # kernels produce and which is held to rtol above.  cuDNN's backward at the P3 shape differs from the CPU
            # implementation by up to ~3e-3 relL2 on identical inputs (tests/tools/diag_grad2.py: dfin agrees to 1e-6 while
            # dfeat_l0 does not), so those torch-only tensors get a looser L2 bound.
            # "dfin" (total) additionally contains head_out's backward-data: one flipped ReLU at a coarse level (24
            # pixels at P6) touches a 3x3 neighbourhood x 256 channels, i.e. a third of the tensor.  The hot path's own
            # contribution is checked strictly through "dfin_direct".
            torch_only = ("dfeat_l" in k) or ("dfin_l" in k) or ("head_in." in k) or ("head_out." in k)
            if device_run and torch_only and rel_l2 <= 3e-2:
                continue
            bad.append("float mismatch %s: max|d|=%.3e scale=%.3e outliers=%.2f%% relL2=%.2e" % (k, err, scale, 100 * frac, rel_l2))
    for k in got:
        if k not in want and not any(s in k for s in skip):
            bad.append("unexpected %s" % k)
    return bad
