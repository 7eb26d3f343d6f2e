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


SAMPLE_LIMIT = {"n": 1024}    # None = keep every element (the bench-size parity tests)


def _sample(t, limit=None):
    limit = SAMPLE_LIMIT["n"] if limit is None else limit
    a = t.detach().float().cpu().reshape(-1).numpy()
    if limit is None or a.size <= limit:
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


def run_case(name, module, impl, device="cpu", boxlist_cls=BoxList, load_fixture=True, prepared=None, tf32=False):
    """impl: 'reference' | 'oracle' | 'product'.  Returns dict[str, np.ndarray].
    prepared: optional (cfg, case, src_feats, src_targets, tgt_feats) instead of build_case(name) (weights already loaded).
    tf32: leave torch's cuDNN TF32 default on (the benchmark's flags) instead of forcing true-fp32 towers."""
    cfg, case, src_feats, src_targets, tgt_feats = prepared if prepared is not None else build_case(name, boxlist_cls)
    if prepared is not None:
        load_fixture = False
    if load_fixture:
        kg, gg = case["fixture"]
        sd = fixture_state_dict(module, seed=99, kernel_gain=kg, gn_gain=gg)
        module.load_state_dict(sd)
    module.to(device) if impl != "reference" else None
    if impl == "product":
        module.record = True      # keep the intermediate results (labels, node rows, DBSCAN masks) of each call in module.last
    if device != "cpu":
        # parity runs compare against an fp32 CPU reference: keep torch's own conv / matmul kernels (the towers,
        # which stay on cuDNN/cuBLAS) in true fp32; the only tf32 arithmetic left is the tcgen05 conditional conv
        torch.backends.cudnn.allow_tf32 = bool(tf32)
        if impl == "product":
            # the product's own tower convolutions (csrc/tower.cu): 3xTF32 (fp32-accurate) for the parity runs, single-pass TF32
            # (= what cuDNN runs under torch's default) at the benchmark's flags
            from scan_b200 import ops as _ops
            _ops.CONV["precise"] = not bool(tf32)
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


TORCH_ONLY = ("dfeat_l", "dfin_l", "head_in.", "head_out.")     # tensors whose values pass through cuDNN's own backward
REPORT = []    # (key, max|d|, scale, outlier fraction, relL2) of every tensor that needed a relaxed rule (printed by the tests)


def is_torch_only(k):
    return any(t in k for t in TORCH_ONLY) and "dfin_direct" not in k


def compare(got, want, rtol=1e-3, atol_scale=1e-3, skip=(), device_run=False, only=None, cudnn_l2=1e-2, flip_frac=0.02, flip_l2=3e-3):
    """Bit-exact for integer results, rtol (relative to the tensor's max magnitude) for floats.
    only: optional predicate on the key (compare a subset).  Returns list of human-readable mismatches.

    Relaxed rules (each use is appended to REPORT):
      * tensors downstream of a tower ReLU (d(features), d(features_in) totals, tower / classifier-hidden gradients) may hold a
        FEW entries whose ReLU mask flipped (pre-activation within ~1e-6 of zero on one side): pass if <= 2 % of the entries
        exceed the bound and the relative L2 error is <= flip_l2 (3e-3).  `dfin_direct` (the hot path's own backward, no tower ReLU in
        it) is NOT in this class: it must meet the max-norm bound, up to a single flipped unit of the node classifier's hidden
        ReLU (<= 0.01 % of the entries beyond the bound AND relative L2 <= 3e-4; logged like every other relaxed rule).
      * device runs against the CPU oracle only: TORCH_ONLY tensors are produced by cuDNN's backward-data / backward-filter,
        whose result differs from the CPU convolution on identical inputs (tests/tools/diag_grad2.py); they get the relative-L2
        bound `cudnn_l2` (measured worst case 6.9e-3 at the P4 shape, round 2 run A; it was 3e-2 in round 1).  The scan_b200
        kernels on that path (GroupNorm+ReLU, add+ReLU, unpack) are held to the 1e-3 max-norm bound by
        `compare(got, twin, only=is_torch_only, ...)` against a twin run whose towers use torch's own GPU ops around the SAME
        cuDNN calls (tests/test_gpu_module.py), so only cuDNN's own difference is excused here."""
    bad = []
    for k, w in want.items():
        if any(s in k for s in skip) or (only is not None and not only(k)):
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
            d = np.abs(g.astype(np.float64) - w.astype(np.float64))
            frac = float((d > rtol * scale).mean())
            rel_l2 = float(np.linalg.norm(d) / max(np.linalg.norm(w.astype(np.float64)), 1e-30))
            relu_path = is_torch_only(k) or ("proto_cls_hidden" in k)
            if relu_path and frac <= flip_frac and rel_l2 <= flip_l2:
                REPORT.append((k, err, scale, frac, rel_l2, "relu-flip rule"))
                continue
            # d(features_in) of the hot path alone passes through no tower ReLU but through the node classifier's hidden ReLU
            # (proto_cls_hidden): ONE unit whose pre-activation is within fp32 rounding of zero flips between two fp32-accurate
            # implementations and changes that node's gradient row.  Ten times tighter than the tower rule on both counts:
            # <= 0.01 % of the entries beyond the bound and relative L2 <= 3e-4 (seen: 1.3e-3 of the maximum on < 0.005 % of a
            # level, relative L2 9e-5).
            if "dfin_direct" in k and frac <= 1e-4 and rel_l2 <= 3e-4:
                REPORT.append((k, err, scale, frac, rel_l2, "classifier-relu-flip rule (tight)"))
                continue
            if device_run and cudnn_l2 and is_torch_only(k) and rel_l2 <= cudnn_l2:
                REPORT.append((k, err, scale, frac, rel_l2, "cudnn-vs-cpu L2 rule"))
                continue
            bad.append("float mismatch %s: max|d|=%.3e scale=%.3e outliers=%.2f%% relL2=%.2e" % (k, err, scale, 100 * frac, rel_l2))
    if only is None:
        for k in got:
            if k not in want and not any(s in k for s in skip):
                bad.append("unexpected %s" % k)
    return bad


# ----------------------------------------------------------------------------------------------------------------------
# Twin run for device tests: the product module with its tower-side kernels (GroupNorm+ReLU, add+ReLU, NCHW<->rows) replaced
# by torch's own GPU ops around the SAME cuDNN convolution calls.  TEST CODE: the product never takes this path.
# ----------------------------------------------------------------------------------------------------------------------
class torch_tower_twin(object):
    """convs="scan": the twin keeps the product's own tower convolutions (csrc/tower.cu, checked against fp64 convolutions in
    tests/test_gpu_kernels.py) and swaps only GroupNorm+ReLU / add+ReLU / pack for torch ops, so the strict bound isolates
    those kernels; convs="cudnn": torch's convolutions as well (two fp32-accurate convolution implementations then differ
    through ReLU flips, like cuDNN against the CPU)."""

    def __init__(self, convs="scan"):
        self.convs = convs

    def __enter__(self):
        import torch.nn.functional as F
        from scan_b200 import ops
        self.ops = ops
        self.saved = (ops.gn_relu_levels, ops.add_relu_levels, ops.pack_levels)

        def gn_relu_levels(geo, gamma, beta, eps, xs, conv_bias=None, stats=None):     # stats (the convolution's by-product) unused
            out = []
            for x in xs:
                if conv_bias is not None:
                    x = x + conv_bias.view(1, -1, 1, 1)
                out.append(torch.relu(F.group_norm(x, 32, gamma, beta, eps)).contiguous(memory_format=torch.channels_last))
            return out

        def add_relu_levels(geo, bias, us, vs=None):
            out = []
            for i, u in enumerate(us):
                y = u if vs is None else u + vs[i]
                if bias is not None:
                    y = y + bias.view(1, -1, 1, 1)
                out.append(torch.relu(y).contiguous(memory_format=torch.channels_last))
            return out

        def pack_levels(geo, feats):
            return [f.contiguous(memory_format=torch.channels_last) for f in feats]

        ops.gn_relu_levels, ops.add_relu_levels, ops.pack_levels = gn_relu_levels, add_relu_levels, pack_levels
        from scan_b200 import condgraph
        self.towers = condgraph.TOWERS
        self.saved_impl = self.towers["impl"]
        self.towers["impl"] = self.convs
        return self

    def __exit__(self, *exc):
        self.ops.gn_relu_levels, self.ops.add_relu_levels, self.ops.pack_levels = self.saved
        self.towers["impl"] = self.saved_impl
        return False


# ----------------------------------------------------------------------------------------------------------------------
# The benchmark's configuration as a parity case (BASELINE.json configs[1]: n source + n target images, 800x1344, K = 9,
# "trained-like" weights: seeded fixture, settle steps that fill the paradigm buffer, closed-form manifestation fit).
# The fit runs ONCE on the CPU oracle; the product loads the resulting state dict, so both see identical bits.
# ----------------------------------------------------------------------------------------------------------------------
def prepare_bench_like(n, settle=6, steps=("source", "target", "eval"), preset="c2f", num_fg=8, **over):
    from oracle.condgraph_oracle import build_oracle
    from scan_b200.fixtures import fit_trained_like
    cfg = scan_cfg(preset, **over)
    src_feats, src_targets = make_workload(n, num_fg, seed=1234, dir_seed=77)
    tgt_feats, _ = make_workload(n, num_fg, seed=4321, dir_seed=77)
    oracle = build_oracle(cfg)
    oracle.load_state_dict(fixture_state_dict(oracle, seed=99))
    oracle.use_sklearn = False          # oracle/dbscan_oracle.c: the point sets exceed what sklearn can hold
    oracle.train()
    with torch.no_grad():
        for _ in range(settle):
            oracle(None, [f[:1] for f in src_feats], targets=src_targets[:1], mode="source")
    fit_trained_like(oracle, [f[:1] for f in src_feats], src_targets[:1])
    case = dict(cfg=(preset, over), n=n, steps=list(steps), fixture=None)
    counter = oracle.counter_rnn.counter if hasattr(oracle, "counter_rnn") else None
    return oracle, (cfg, case, src_feats, src_targets, tgt_feats), counter


def load_like(module, oracle, counter):
    module.load_state_dict({k: v.clone() for k, v in oracle.state_dict().items()})
    if counter is not None:
        module.counter_rnn.counter = counter
    return module
