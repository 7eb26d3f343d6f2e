"""CPU ORACLE for the condgraph middle head -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch torch-CPU / numpy restatement of the reference's algorithm for the hot path
(SURVEY.md §8a rows a1-a17).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; `scan_b200/` never does.

Parity status: PINNED AGAINST THE REFERENCE ITSELF.  The reference's own tests hold no
golden vector for this path, so the pin is (1) `tests/test_oracle_vs_reference.py`,
which runs the unmodified reference (imported through `oracle/ref_shim.py`, only possible in
the build container) and this file on the same seeded inputs, and (2) the committed fixtures
under `tests/golden/` that `tests/tools/make_golden.py` generated from the unmodified reference.

Every function cites the reference file:line it follows (paths relative to
/root/reference/fcos_core/).  Third-party arithmetic on the path: `sklearn.cluster.DBSCAN`
(call site modeling/rpn/fcos/loss.py:416; sklearn 1.9.0 in this image, un-pinned by the
reference) -- restated below in `dbscan_labels` and in `oracle/dbscan_oracle.c`;
`numpy.linspace` (loss.py:448, 503) -- restated in `floor_linspace`.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

INF_AREA = 100000000.0  # loss.py:22
SIZES_OF_INTEREST = [(-1.0, 64.0), (64.0, 128.0), (128.0, 256.0), (256.0, 512.0), (512.0, INF_AREA)]  # loss.py:263-269


# --------------------------------------------------------------------------------------
# a2 / a3: locations and FCOS ground-truth assignment
# --------------------------------------------------------------------------------------
def level_locations(h, w, stride):
    """condgraph.py:642-655: (x*s + s//2, y*s + s//2), y-major then x."""
    ys = torch.arange(h, dtype=torch.float32) * stride + stride // 2
    xs = torch.arange(w, dtype=torch.float32) * stride + stride // 2
    return xs.repeat(h), ys.repeat_interleave(w)


def fcos_assign(level_shapes, strides, boxes_per_image, labels_per_image, return_reg=False):
    """loss.py:262-343.  Returns a list (one per level) of int64 label vectors laid out
    image-major ([N*H_l*W_l], loss.py:289-294).  fp32 arithmetic; ranges inclusive on both
    ends; area uses the +1 convention (structures/bounding_box.py:229-230); ties resolved to
    the first minimum (torch.min); a location with no admissible box gets label 0."""
    out = []
    regs = []
    for (h, w), s, (lo, hi) in zip(level_shapes, strides, SIZES_OF_INTEREST):
        xs, ys = level_locations(h, w, s)
        per_image = []
        per_image_reg = []
        for boxes, labels in zip(boxes_per_image, labels_per_image):
            boxes = boxes.float()
            xs, ys = xs.to(boxes.device), ys.to(boxes.device)
            if boxes.shape[0] == 0:
                raise RuntimeError("G == 0 is unsupported by the reference (empty min, loss.py:333)")
            area = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)
            l = xs[:, None] - boxes[None, :, 0]
            t = ys[:, None] - boxes[None, :, 1]
            r = boxes[None, :, 2] - xs[:, None]
            b = boxes[None, :, 3] - ys[:, None]
            mn = torch.minimum(torch.minimum(l, t), torch.minimum(r, b))
            mx = torch.maximum(torch.maximum(l, t), torch.maximum(r, b))
            ok = (mn > 0) & (mx >= lo) & (mx <= hi)
            cost = torch.where(ok, area[None, :].expand_as(mn), torch.full_like(mn, INF_AREA))
            best, idx = cost.min(dim=1)
            lab = labels.long().to(boxes.device)[idx]
            lab = torch.where(best == INF_AREA, torch.zeros_like(lab), lab)
            per_image.append(lab)
            # loss.py:111: regression targets (l, t, r, b) of the chosen box (index 0 where nothing matched)
            per_image_reg.append(torch.stack([l, t, r, b], dim=2)[torch.arange(l.shape[0], device=l.device), idx])
        out.append(torch.cat(per_image))
        regs.append(torch.cat(per_image_reg))
    if return_reg:
        return out, regs
    return out


# --------------------------------------------------------------------------------------
# numpy.linspace restatement (loss.py:448, 503)
# --------------------------------------------------------------------------------------
def floor_linspace(stop, num):
    """floor(np.linspace(0, stop, num)).astype(int) without numpy's helper:
    y_i = float64(i) * (stop/(num-1)), last element forced to `stop` (numpy 2.3 semantics)."""
    if num <= 0:
        return np.zeros((0,), dtype=np.int64)
    if num == 1:
        return np.zeros((1,), dtype=np.int64)
    step = float(stop) / float(num - 1)
    if step == 0.0:
        y = np.zeros(num, dtype=np.float64)
    else:
        y = np.arange(num, dtype=np.float64) * step
    y[-1] = float(stop)
    return np.floor(y).astype(np.int64)


def nhwc_rows(feature):
    """features[l].permute(0,2,3,1).reshape(-1,C) (loss.py:440)."""
    n, c, h, w = feature.shape
    return feature.permute(0, 2, 3, 1).reshape(-1, c)


# --------------------------------------------------------------------------------------
# a4: source-domain node sampling
# --------------------------------------------------------------------------------------
def source_node_indices(labels_per_level, with_bg=True):
    """loss.py:430-458.  Returns (level ids, row ids within the level's NHWC row space, labels)
    in the reference's node order: [neg(P3..P7), pos(P3..P7)]."""
    neg, pos = [], []
    for l, lab in enumerate(labels_per_level):
        lab = lab.reshape(-1)
        p = torch.nonzero(lab > 0).reshape(-1)
        n = torch.nonzero(lab == 0).reshape(-1)
        pos.append((l, p, lab[p]))
        if with_bg:
            if p.numel() > n.numel():
                sel = n
            else:
                k = floor_linspace(n.numel() - 2, p.numel())
                sel = n[torch.from_numpy(k).to(n.device)] if k.size else n[:0]   # negative k wraps like python indexing
            neg.append((l, sel, torch.zeros_like(sel)))
    seq = (neg + pos) if with_bg else pos
    lv = torch.cat([torch.full_like(idx, l) for l, idx, _ in seq])
    rows = torch.cat([idx for _, idx, _ in seq])
    labs = torch.cat([lb for _, _, lb in seq])
    return lv, rows, labs


def gather_nodes(features, lv, rows):
    rows_per_level = [nhwc_rows(f) for f in features]
    chunks = []
    # node order is a concatenation of per-level runs, so gather run by run
    start = 0
    lv_np = lv.cpu().numpy()
    while start < len(lv_np):
        end = start
        while end < len(lv_np) and lv_np[end] == lv_np[start]:
            end += 1
        chunks.append(rows_per_level[int(lv_np[start])][rows[start:end]])
        start = end
    return torch.cat(chunks, dim=0) if chunks else rows_per_level[0][:0]


# --------------------------------------------------------------------------------------
# a13: DBSCAN target-domain sampling
# --------------------------------------------------------------------------------------
_C_DBSCAN = None


def _c_dbscan():
    """oracle/dbscan_oracle.c, when `make -C oracle` has been run (the same restatement in plain C, OpenMP)."""
    global _C_DBSCAN
    if _C_DBSCAN is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbscan_oracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.dbscan_oracle.restype = ctypes.c_int
            lib.dbscan_oracle.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_int,
                                          ctypes.c_void_p]
            _C_DBSCAN = lib
        else:
            _C_DBSCAN = False
    return _C_DBSCAN


def dbscan_labels_c(points, eps, min_samples=5):
    lib = _c_dbscan()
    if not lib:
        raise RuntimeError("oracle/libdbscan_oracle.so not built (make -C oracle)")
    x = np.ascontiguousarray(points, dtype=np.float32)
    labels = np.empty(x.shape[0], dtype=np.int32)
    rc = lib.dbscan_oracle(x.ctypes.data, x.shape[0], x.shape[1], float(eps), int(min_samples), labels.ctypes.data)
    if rc != 0:
        raise MemoryError("dbscan_oracle failed")
    return labels.astype(np.int64)


def dbscan_labels(points, eps, min_samples=5, block=2048):
    """sklearn.cluster.DBSCAN(eps, min_samples=5, euclidean).fit_predict restated
    (sklearn 1.9.0: neighbors/_base.py brute radius query with float64 accumulation,
    `d2 <= eps^2` including self; cluster/_dbscan_inner.pyx:17-43 index-order expansion).
    Equivalent deterministic form: core = |N_eps| >= min_samples; components of the
    core-core graph; cluster id = rank of the component's smallest index; a border point takes
    the id of the FIRST cluster (in index order of expansion) that reaches it == the minimum
    cluster id among its core neighbours; others -1.  O(n*block) memory."""
    x = np.ascontiguousarray(points, dtype=np.float64)
    n = x.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    sq = np.einsum("ij,ij->i", x, x)
    eps2 = float(eps) * float(eps)
    counts = np.zeros(n, dtype=np.int64)
    for s in range(0, n, block):
        d2 = sq[s:s + block, None] + sq[None, :] - 2.0 * (x[s:s + block] @ x.T)
        np.maximum(d2, 0.0, out=d2)
        counts[s:s + block] = (d2 <= eps2).sum(1)
    core = counts >= min_samples
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    core_idx = np.nonzero(core)[0]
    xc = x[core_idx]
    sqc = sq[core_idx]
    for s in range(0, len(core_idx), block):
        d2 = sqc[s:s + block, None] + sqc[None, :] - 2.0 * (xc[s:s + block] @ xc.T)
        np.maximum(d2, 0.0, out=d2)
        ii, jj = np.nonzero(d2 <= eps2)
        for a, b in zip(core_idx[ii + s], core_idx[jj]):
            if a < b:
                ra, rb = find(a), find(b)
                if ra != rb:
                    if ra < rb:
                        parent[rb] = ra
                    else:
                        parent[ra] = rb
    roots = np.array([find(i) for i in core_idx], dtype=np.int64)
    uniq = np.unique(roots)  # sorted: root == smallest index of its component
    rank = {int(r): k for k, r in enumerate(uniq)}
    labels = np.full(n, -1, dtype=np.int64)
    for i, r in zip(core_idx, roots):
        labels[i] = rank[int(r)]
    border = np.nonzero(~core)[0]
    if len(border) and len(core_idx):
        core_lab = labels[core_idx]
        for s in range(0, len(border), block):
            b = border[s:s + block]
            d2 = sq[b, None] + sqc[None, :] - 2.0 * (x[b] @ xc.T)
            np.maximum(d2, 0.0, out=d2)
            near = d2 <= eps2
            cand = np.where(near, core_lab[None, :], np.iinfo(np.int64).max)
            best = cand.min(1)
            has = near.any(1)
            labels[b[has]] = best[has]
    return labels


def dbscan_points(act_fg, feature, thr):
    """loss.py:397-414 without materialising the [CLS,N,C,H,W] tensor: selected flat
    indices ((n*CLS+cls)*H+y)*W+x in ascending order and their points feat[n,:,y,x]*act."""
    n, cls, h, w = act_fg.shape
    mask = (act_fg > thr).reshape(-1)
    flat = torch.nonzero(mask).reshape(-1)
    x_ = flat % w
    y_ = (flat // w) % h
    c_ = (flat // (w * h)) % cls
    n_ = flat // (w * h * cls)
    pts = feature[n_, :, y_, x_] * act_fg[n_, c_, y_, x_][:, None]
    return flat, pts


def dbscan_location_mask(act_fg, feature, eps, thr, use_sklearn=True):
    """DBSCAN_batch_cpu (loss.py:397-423): boolean positive mask over the N*H*W locations."""
    n, cls, h, w = act_fg.shape
    act_fg = act_fg.detach()
    feature = feature.detach()
    flat, pts = dbscan_points(act_fg, feature, thr)
    dev = act_fg.device      # on CUDA this is the reference's own D2H -> DBSCAN -> H2D round trip (loss.py:414-421)
    val = torch.zeros(n * cls * h * w, device=dev)
    val[flat] = 1.0
    labels = None
    if pts.numel() and bool(pts.bool().any()):
        if use_sklearn:
            from sklearn.cluster import DBSCAN
            labels = DBSCAN(eps=eps, n_jobs=-1).fit_predict(pts.cpu().numpy())
        elif _c_dbscan():
            labels = dbscan_labels_c(pts.cpu().numpy(), eps)
        else:
            labels = dbscan_labels(pts.cpu().numpy(), eps)
        y = labels.copy()
        y[y < 0] = 1                       # loss.py:417: noise kept as "1", cluster 0 dropped
        val[flat] = torch.from_numpy(y.astype(np.float32)).to(dev)
    pos = val.reshape(n, cls, h, w).permute(0, 2, 3, 1).reshape(-1, cls).sum(-1).bool()
    return pos, flat, labels


def target_node_indices(act_maps, features, eps, thr, use_sklearn=True, sampling="dbscan", plabel_th=0.9):
    """loss.py:464-518.  Returns (lv, rows, plabels, per-level masks); the first three are None
    when no level has a positive location."""
    neg, pos = [], []
    masks = []
    for l, (act, feat) in enumerate(zip(act_maps, features)):
        k = act.shape[1]
        flat_act = act.permute(0, 2, 3, 1).reshape(-1, k)
        if sampling == "dbscan":
            conf, _, _ = dbscan_location_mask(act[:, 1:], feat, eps, thr, use_sklearn)
        elif sampling == "score_threshold":
            conf = (flat_act[:, 1:] > plabel_th).sum(-1).bool()      # loss.py:479-481
        else:
            raise KeyError("unknown target labels!")
        masks.append(conf)
        if bool(conf.any()):
            p = torch.nonzero(conf).reshape(-1)
            n_ = torch.nonzero(~conf).reshape(-1)
            plab = flat_act[p, 1:].argmax(-1) + 1
            kidx = floor_linspace(n_.numel() - 2, p.numel())
            if n_.numel() == 0:
                raise IndexError("no negative location left at level %d (loss.py:503-504)" % l)
            sel = n_[torch.from_numpy(kidx).to(n_.device)]
            pos.append((l, p, plab))
            neg.append((l, sel, torch.zeros_like(sel)))
    if not pos:
        return None, None, None, masks
    seq = neg + pos
    lv = torch.cat([torch.full_like(idx, l) for l, idx, _ in seq])
    rows = torch.cat([idx for _, idx, _ in seq])
    labs = torch.cat([lb for _, _, lb in seq])
    return lv, rows, labs, masks


# --------------------------------------------------------------------------------------
# a5: the hand-rolled multi-head attention (layers/transformer.py:36-90)
# --------------------------------------------------------------------------------------
def chunked_attention(x, p, heads=4, drop_attn=None, drop_out=None):
    """x [M,256].  The reference's `.view(B*heads, -1, d)` on a [1,M,256] tensor splits the
    row-major [M*4,64] sub-token matrix into `heads` consecutive chunks of M sub-tokens each;
    scale = (64 // 4) ** -0.5 = 0.25 (transformer.py:75).  drop_* are optional pre-drawn masks
    already divided by keep-prob."""
    m, c = x.shape
    d = c // heads
    q = F.linear(x, p["linear_q.weight"], p["linear_q.bias"]).reshape(heads, m, d)
    k = F.linear(x, p["linear_k.weight"], p["linear_k.bias"]).reshape(heads, m, d)
    v = F.linear(x, p["linear_v.weight"], p["linear_v.bias"]).reshape(heads, m, d)
    scale = float((d // heads) ** -0.5)
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * scale, dim=2)
    if drop_attn is not None:
        att = att * drop_attn
    ctx = torch.bmm(att, v).reshape(m, c)
    out = F.linear(ctx, p["linear_final.weight"], p["linear_final.bias"])
    if drop_out is not None:
        out = out * drop_out
    return F.layer_norm(x + out, (c,), p["layer_norm.weight"], p["layer_norm.bias"], 1e-5)


# --------------------------------------------------------------------------------------
# a6: per-class GCN (condgraph.py:262-302)
# --------------------------------------------------------------------------------------
def cosine_matrix(a, b, eps=1e-8):
    """condgraph.py:35-43."""
    an = a / a.norm(dim=1, keepdim=True).clamp(min=eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp(min=eps)
    return an @ bn.t()


def gcn_edge(nodes, norm):
    if norm == "NO":
        return (nodes @ nodes.t()).softmax(-1).detach()
    if norm == "cosine_detached":
        return cosine_matrix(nodes, nodes).softmax(-1).detach()
    raise AttributeError("edge norm %r needs edge_project_u/v which the reference never defines" % norm)


def gcn_forward(nodes, adj, w1, b1, w2, b2, out_act="relu", shortcut=False):
    x = torch.relu(F.linear(adj @ nodes, w1, b1))
    y = F.linear(adj @ x, w2, b2)
    if out_act == "softmax":
        y = y.softmax(-1)
    elif out_act == "sigmoid":
        y = y.sigmoid()
    elif out_act == "tanh":
        y = y.tanh()
    elif out_act == "relu":
        y = torch.relu(y)
    elif out_act != "NO":
        raise KeyError("unknown gcn output activation")
    return y + nodes if shortcut else y


# --------------------------------------------------------------------------------------
# a11 / a17: losses
# --------------------------------------------------------------------------------------
def softmax_focal_loss(logits, targets, gamma=2.0):
    """layers/sigmoid_focal_loss_wbg.py:7-64 (FocalLoss, alpha=1, mean)."""
    p = logits.softmax(dim=1)
    pt = p.gather(1, targets.view(-1, 1))
    if bool((pt < 1e-15).any()):
        pt = pt.clamp(min=1e-15)
    return (-(1 - pt).pow(gamma) * pt.log()).mean()


def bce_focal_loss(logits, target, gamma=2.0, alpha=0.25):
    """layers/sigmoid_focal_loss_wbg.py:148-177 (BCEFocalLoss, elementwise mean)."""
    pt = torch.sigmoid(logits).clamp(min=0.00001).clamp(max=0.99999)
    loss = -alpha * (1 - pt) ** gamma * target * torch.log(pt) \
           - (1 - alpha) * pt ** gamma * (1 - target) * torch.log(1 - pt)
    return loss.mean()


def sigmoid_focal_loss_elementwise(logits, targets, gamma, alpha):
    """csrc/cuda/SigmoidFocalLoss_cuda.cu:21-58 == layers/sigmoid_focal_loss.py:40-53.
    logits [R,C] fp32, targets [R] int (class ids 1..C, 0 = background, <0 ignored)."""
    c = logits.shape[1]
    cls = torch.arange(1, c + 1, dtype=targets.dtype).unsqueeze(0)
    t = targets.unsqueeze(1)
    p = torch.sigmoid(logits)
    # SigmoidFocalLoss_cuda.cu:45: logf(max(p, FLT_MIN)) -- the CUDA kernel (the one the reference trains with) clamps, the
    # CPU mirror (layers/sigmoid_focal_loss.py:50) does not; they differ only where sigmoid(x) is subnormal (x < -87.3)
    term1 = (1 - p) ** gamma * torch.log(p.clamp(min=1.17549435e-38))
    # CUDA kernel uses the stable form of log(1-p): -x*(x>=0) - log(1+exp(x-2x*(x>=0)))
    xs = logits
    log1mp = -xs * (xs >= 0).float() - torch.log1p(torch.exp(xs - 2 * xs * (xs >= 0).float()))
    term2 = p ** gamma * log1mp
    return -(t == cls).float() * term1 * alpha - ((t != cls) & (t >= 0)).float() * term2 * (1 - alpha)


def sigmoid_focal_loss_backward_elementwise(logits, targets, d_losses, gamma, alpha):
    """csrc/cuda/SigmoidFocalLoss_cuda.cu:61-101: the reference's ANALYTIC backward (what `_C.sigmoid_focalloss_backward`
    returns).  It differentiates log(p) without the FLT_MIN clamp of the forward, so it differs from autograd of
    `sigmoid_focal_loss_elementwise` where sigmoid(x) is subnormal (x < -87.3); everywhere else the two agree."""
    c = logits.shape[1]
    cls = torch.arange(1, c + 1, dtype=targets.dtype).unsqueeze(0)
    t = targets.unsqueeze(1)
    x = logits
    p = torch.sigmoid(x)
    ge = (x >= 0).float()
    term1 = (1 - p) ** gamma * (1 - p - p * gamma * torch.log(p.clamp(min=1.17549435e-38)))
    term2 = p ** gamma * ((-x * ge - torch.log1p(torch.exp(x - 2 * x * ge))) * (1 - p) * gamma - p)
    return (-(t == cls).float() * term1 * alpha - ((t != cls) & (t >= 0)).float() * term2 * (1 - alpha)) * d_losses


def fcos_loss_computation(level_shapes, strides, boxes_per_image, labels_per_image, box_cls, box_regression, centerness,
                          gamma=2.0, alpha=0.25):
    """FCOSLossComputation.__call__ (loss.py:168-230) with IOULoss (layers/iou_loss.py:5-38), compute_centerness_targets
    (loss.py:128-133) and nn.BCEWithLogitsLoss: returns (cls_loss, reg_loss, centerness_loss)."""
    n = box_cls[0].shape[0]
    c = box_cls[0].shape[1]
    labels, regs = fcos_assign(level_shapes, strides, boxes_per_image, labels_per_image, return_reg=True)
    cls_flat = torch.cat([x.permute(0, 2, 3, 1).reshape(-1, c) for x in box_cls])
    reg_flat = torch.cat([x.permute(0, 2, 3, 1).reshape(-1, 4) for x in box_regression])
    ctr_flat = torch.cat([x.reshape(-1) for x in centerness])
    lab = torch.cat(labels)
    tgt = torch.cat(regs)
    pos = torch.nonzero(lab > 0).squeeze(1)
    cls_loss = sigmoid_focal_loss_elementwise(cls_flat, lab.int(), gamma, alpha).sum() / (pos.numel() + n)
    reg_p, reg_t, ctr_p = reg_flat[pos], tgt[pos], ctr_flat[pos]
    if pos.numel() == 0:
        return cls_loss, reg_p.sum(), ctr_p.sum()
    lr, tb = reg_t[:, [0, 2]], reg_t[:, [1, 3]]
    ctr_t = torch.sqrt((lr.min(-1)[0] / lr.max(-1)[0]) * (tb.min(-1)[0] / tb.max(-1)[0]))
    t_area = (reg_t[:, 0] + reg_t[:, 2]) * (reg_t[:, 1] + reg_t[:, 3])
    p_area = (reg_p[:, 0] + reg_p[:, 2]) * (reg_p[:, 1] + reg_p[:, 3])
    w_i = torch.min(reg_p[:, 0], reg_t[:, 0]) + torch.min(reg_p[:, 2], reg_t[:, 2])
    h_i = torch.min(reg_p[:, 3], reg_t[:, 3]) + torch.min(reg_p[:, 1], reg_t[:, 1])
    inter = w_i * h_i
    union = t_area + p_area - inter
    iou = -torch.log((inter + 1.0) / (union + 1.0))
    reg_loss = (iou * ctr_t).sum() / ctr_t.sum() if bool(ctr_t.sum() > 0) else iou.mean()
    ctr_loss = F.binary_cross_entropy_with_logits(ctr_p, ctr_t)
    return cls_loss, reg_loss, ctr_loss


def ensemble(mode, cls_logits, act_maps):
    """fcos.py:162-169 followed by the sigmoid of inference.py:68 (common mode only):
    returns the per-level class PROBABILITY maps the post-processor thresholds."""
    out = []
    for i, act in enumerate(act_maps):
        if mode == "light":
            out.append(act[:, 1:])
        elif mode == "precision":
            out.append(0.5 * cls_logits[i].sigmoid() + 0.5 * act[:, 1:])
        else:
            out.append(cls_logits[i].sigmoid())
    return out


# --------------------------------------------------------------------------------------
# f4: FCOS post-processor (inference.py:54-194)
# --------------------------------------------------------------------------------------
def nms_indices(boxes, scores, thresh):
    """The reference's `_C.nms` on a GPU (csrc/cuda/nms.cu:13-129): greedy in score-descending order, a box is suppressed
    when IoU (with the +1 pixel convention) is STRICTLY greater than `thresh`; returns the kept ORIGINAL indices in
    ascending order (nms.cu:123-129 sorts them).  (The CPU twin csrc/cpu/nms_cpu.cpp uses >= and is not what inference runs.)"""
    b = boxes.detach().cpu().double().numpy() if torch.is_tensor(boxes) else np.asarray(boxes, dtype=np.float64)
    sc = scores.detach().cpu().numpy() if torch.is_tensor(scores) else np.asarray(scores)
    n = b.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.int64)
    b32 = b.astype(np.float32)
    order = np.argsort(-sc, kind="stable")
    area = (b32[:, 2] - b32[:, 0] + np.float32(1)) * (b32[:, 3] - b32[:, 1] + np.float32(1))
    removed = np.zeros(n, dtype=bool)
    keep = []
    for oi, i in enumerate(order):
        if removed[i]:
            continue
        keep.append(i)
        rest = order[oi + 1:]
        left = np.maximum(b32[i, 0], b32[rest, 0])
        right = np.minimum(b32[i, 2], b32[rest, 2])
        top = np.maximum(b32[i, 1], b32[rest, 1])
        bottom = np.minimum(b32[i, 3], b32[rest, 3])
        w = np.maximum(right - left + np.float32(1), np.float32(0))
        h = np.maximum(bottom - top + np.float32(1), np.float32(0))
        inter = w * h
        iou = inter / (area[i] + area[rest] - inter)
        removed[rest[iou > np.float32(thresh)]] = True
    return torch.from_numpy(np.sort(np.asarray(keep, dtype=np.int64)))


def fcos_postprocess(level_shapes, strides, probs, box_regression, centerness, image_sizes, pre_nms_thresh, pre_nms_top_n,
                     nms_thresh, post_top_n, min_size=0):
    """FCOSPostProcessor.forward (inference.py:54-194) on per-level class probability maps: forward_for_single_feature_map
    (:54-121), cat over levels, select_over_all_levels (:148-194) with boxlist_nms (structures/boxlist_ops.py:9-31),
    clip_to_image (structures/bounding_box.py:214-224) and remove_small_boxes (boxlist_ops.py:58-74).
    Returns per image (boxes [D,4], scores [D], labels [D]).  When a level has more candidates than pre_nms_top_n the reference's
    `topk(sorted=False)` leaves their order unspecified; this restatement keeps candidate order."""
    n_img = probs[0].shape[0]
    per_image = [[] for _ in range(n_img)]
    for (h, w), s, p, r, c in zip(level_shapes, strides, probs, box_regression, centerness):
        n, cdim = p.shape[0], p.shape[1]
        xs, ys = level_locations(h, w, s)
        box_cls = p.permute(0, 2, 3, 1).reshape(n, -1, cdim)
        reg = r.permute(0, 2, 3, 1).reshape(n, -1, 4)
        ctr = c.permute(0, 2, 3, 1).reshape(n, -1).sigmoid()
        cand = box_cls > pre_nms_thresh
        top_n = cand.reshape(n, -1).sum(1).clamp(max=pre_nms_top_n)
        box_cls = box_cls * ctr[:, :, None]
        for i in range(n):
            nz = cand[i].nonzero()
            loc, cls = nz[:, 0], nz[:, 1] + 1
            sc = box_cls[i][cand[i]]
            if int(cand[i].sum()) > int(top_n[i]):
                _, idx = sc.topk(int(top_n[i]), sorted=False)
                idx = idx.sort().values                        # candidate order (see docstring)
                sc, loc, cls = sc[idx], loc[idx], cls[idx]
            rg = reg[i][loc]
            det = torch.stack([xs[loc] - rg[:, 0], ys[loc] - rg[:, 1], xs[loc] + rg[:, 2], ys[loc] + rg[:, 3]], dim=1)
            ih, iw = image_sizes[i]
            det[:, 0].clamp_(min=0, max=iw - 1)
            det[:, 1].clamp_(min=0, max=ih - 1)
            det[:, 2].clamp_(min=0, max=iw - 1)
            det[:, 3].clamp_(min=0, max=ih - 1)
            keep = ((det[:, 2] - det[:, 0] + 1 >= min_size) & (det[:, 3] - det[:, 1] + 1 >= min_size)).nonzero().squeeze(1)
            per_image[i].append((det[keep], torch.sqrt(sc)[keep], cls[keep]))
    out = []
    for i in range(n_img):
        boxes = torch.cat([t[0] for t in per_image[i]])
        scores = torch.cat([t[1] for t in per_image[i]])
        labels = torch.cat([t[2] for t in per_image[i]])
        rb, rs, rl = [], [], []
        for j in range(1, probs[0].shape[1] + 1):
            inds = (labels == j).nonzero().view(-1)
            keep = nms_indices(boxes[inds], scores[inds], nms_thresh)
            rb.append(boxes[inds][keep])
            rs.append(scores[inds][keep])
            rl.append(torch.full((keep.numel(),), j, dtype=torch.int64))
        boxes, scores, labels = torch.cat(rb), torch.cat(rs), torch.cat(rl)
        if scores.numel() > post_top_n > 0:
            thr, _ = torch.kthvalue(scores, scores.numel() - post_top_n + 1)
            k = (scores >= thr.item()).nonzero().squeeze(1)
            boxes, scores, labels = boxes[k], scores[k], labels[k]
        out.append((boxes, scores, labels))
    return out


# --------------------------------------------------------------------------------------
# The module
# --------------------------------------------------------------------------------------
class _Counter(object):
    """condgraph.py:46-65."""

    def __init__(self, cycle=3, stop=False):
        self.cycle, self.stop, self.counter = cycle, stop, -1

    def __call__(self):
        if self.stop:
            if self.counter != self.cycle:
                self.counter += 1
            return self.counter
        self.counter += 1
        if self.counter == self.cycle:
            self.counter = 0
        return self.counter


def _tower(n_convs, cin, cout, norm):
    layers = []
    for _ in range(n_convs):
        conv = nn.Conv2d(cin, cout, 3, 1, 1)
        nn.init.normal_(conv.weight, std=0.01)
        nn.init.constant_(conv.bias, 0)
        layers.append(conv)
        if norm == "GN":
            layers.append(nn.GroupNorm(32, cin))
        elif norm == "IN":
            layers.append(nn.InstanceNorm2d(cin))
        elif norm == "BN":
            layers.append(nn.BatchNorm2d(cin))
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class _Head(nn.Module):
    def __init__(self, seq):
        super().__init__()
        self.middle_tower = seq

    def forward(self, xs):
        return [self.middle_tower(x) for x in xs]


class _MHA(nn.Module):
    def __init__(self, dim=256, heads=4, dropout=0.1):
        super().__init__()
        self.heads = heads
        self.linear_k = nn.Linear(dim, dim)
        self.linear_v = nn.Linear(dim, dim)
        self.linear_q = nn.Linear(dim, dim)
        self.linear_final = nn.Linear(dim, dim)
        self.layer_norm = nn.LayerNorm(dim)
        self.p_drop = dropout


class OracleCondGraph(nn.Module):
    """Restatement of GRAPHModule (condgraph.py:122-669) with the same state_dict keys."""

    def __init__(self, cfg, in_channels=256):
        super().__init__()
        mh = cfg.MODEL.MIDDLE_HEAD
        self.mh = mh
        self.cfg = cfg
        self.strides = list(cfg.MODEL.FCOS.FPN_STRIDES)
        self.num_fg = cfg.MODEL.FCOS.NUM_CLASSES - 1
        self.with_bg = bool(mh.PROTO_WITH_BG)
        self.K = self.num_fg + int(self.with_bg)
        self.P = mh.PROTO_ITER
        C = mh.PROTO_CHANNEL
        hid = mh.COND_HIDDEN_CHANNEL
        self.head_in = _Head(_tower(mh.NUM_CONVS_IN, in_channels, in_channels, mh.IN_NORM))
        if self.P == 1:
            self.register_buffer("prototype", torch.randn(self.K, C))
        else:
            self.register_buffer("prototype", torch.randn(self.K, C, self.P))
        if mh.CAT_ACT_MAP:
            self.head_out = _Head(_tower(mh.NUM_CONVS_OUT, in_channels + self.K, in_channels, None))
        self.proto_cls_hidden = nn.Linear(mh.GCN2_OUT_CHANNEL, 512)
        self.proto_cls = nn.Linear(512, self.K)
        if mh.GLOBAL_GCN:
            self.multihead_attn = _MHA(256, 4, 0.1)
        else:
            self.gcn_layer1 = nn.Linear(256, mh.GCN1_OUT_CHANNEL)
            self.gcn_layer2 = nn.Linear(mh.GCN1_OUT_CHANNEL, mh.GCN2_OUT_CHANNEL)
            for l in (self.gcn_layer1, self.gcn_layer2):
                nn.init.normal_(l.weight, std=0.01)
                nn.init.constant_(l.bias, 0)
        if mh.USE_RNN:
            self.cond_nx1 = nn.Conv2d(512, 256, kernel_size=(self.P, 1))
            self.cond_rnn = nn.RNN(256, 512, 2, nonlinearity="tanh")
            self.counter_rnn = _Counter(self.P, stop=True)
        elif self.P > 1:
            self.counter = _Counter(self.P)
            self.cond_nx1 = nn.Conv2d(C, hid, kernel_size=(self.P, 1))
            nn.init.normal_(self.cond_nx1.weight)
            nn.init.constant_(self.cond_nx1.bias, 0)
            self.cond_nx1_norm = nn.GroupNorm(32, hid)
        else:
            self.cond_1 = nn.Linear(C, hid)
            nn.init.normal_(self.cond_1.weight, std=0.01)
            nn.init.constant_(self.cond_1.bias, 0)
        self.cond_2 = nn.Linear(hid, 256 + int(bool(mh.COND_WITH_BIAS)))
        for l in (self.cond_2, self.proto_cls, self.proto_cls_hidden):
            nn.init.normal_(l.weight, std=0.01)
            nn.init.constant_(l.bias, 0)
        self.use_sklearn = True
        self.last = {}   # intermediate results exposed for parity tests

    # ---- manifestation (a9; condgraph.py:313-336) -------------------------------------
    def conded_weight(self):
        mh = self.mh
        if mh.USE_RNN:
            seq = self.prototype.permute(2, 0, 1)                      # [P,K,256]
            h = seq
            for layer in range(2):
                wi = getattr(self.cond_rnn, "weight_ih_l%d" % layer)
                wh = getattr(self.cond_rnn, "weight_hh_l%d" % layer)
                bi = getattr(self.cond_rnn, "bias_ih_l%d" % layer)
                bh = getattr(self.cond_rnn, "bias_hh_l%d" % layer)
                state = h.new_zeros(h.shape[1], wh.shape[0])
                outs = []
                for t in range(h.shape[0]):
                    state = torch.tanh(F.linear(h[t], wi, bi) + F.linear(state, wh, bh))
                    outs.append(state)
                h = torch.stack(outs)                                  # [P,K,512]
            # cond_nx1: Conv2d(512->256, kernel (P,1)) over input [K,512,P,1]
            w = self.cond_nx1.weight[:, :, :, 0]                       # [256,512,P]
            return torch.einsum("pkc,ocp->ko", h, w) + self.cond_nx1.bias
        if self.P > 1:
            w = self.cond_nx1.weight[:, :, :, 0]                       # [hid,256,P]
            hcat = torch.einsum("kcp,ocp->ko", self.prototype, w) + self.cond_nx1.bias
            hcat = F.group_norm(hcat, 32, self.cond_nx1_norm.weight, self.cond_nx1_norm.bias, 1e-5)
            return self.cond_2(torch.relu(hcat))
        return self.cond_2(torch.relu(self.cond_1(self.prototype)))

    # ---- conditional conv (a10; condgraph.py:619-629) ---------------------------------
    def dynamic_conv(self, feat, kernel_par):
        if self.mh.COND_WITH_BIAS:
            return F.conv2d(feat, kernel_par[:, :-1].reshape(self.K, -1, 1, 1), bias=kernel_par[:, -1])
        return F.conv2d(feat, kernel_par.reshape(self.K, -1, 1, 1))

    def act_of(self, logits):
        return logits.softmax(dim=1) if self.mh.ACT_LOSS == "softmaxFL" else logits.sigmoid()

    # ---- graph aggregation (a5-a7; condgraph.py:386-421) ------------------------------
    def forward_gcns(self, nodes, labels):
        mh = self.mh
        proto = nodes.new_zeros(self.K, nodes.shape[1])
        shift = 0 if self.with_bg else 1
        if mh.GLOBAL_GCN:
            a = self.multihead_attn
            params = {k: v for k, v in a.named_parameters()}
            da = do = None
            if self.training and a.p_drop > 0:
                m = nodes.shape[0]
                keep = 1.0 - a.p_drop
                da = torch.bernoulli(torch.full((a.heads, m, m), keep, device=nodes.device)) / keep
                do = torch.bernoulli(torch.full((m, nodes.shape[1]), keep, device=nodes.device)) / keep
            out = chunked_attention(nodes, params, a.heads, da, do)
            if mh.GCN_SHORTCUT:
                out = out + nodes
        else:
            out = nodes.clone()
            for i in range(self.K):
                idx = labels == i + shift
                if bool(idx.any()):
                    sub = nodes[idx]
                    adj = gcn_edge(sub, mh.GCN_EDGE_NORM)
                    out[idx] = gcn_forward(sub, adj, self.gcn_layer1.weight, self.gcn_layer1.bias,
                                           self.gcn_layer2.weight, self.gcn_layer2.bias,
                                           mh.GCN_OUT_ACTIVATION, mh.GCN_SHORTCUT)
        for i in range(self.K):
            idx = labels == i + shift
            if bool(idx.any()):
                proto[i] = out[idx].mean(0)
        logits = self.proto_cls(torch.relu(self.proto_cls_hidden(out)))
        node_loss = mh.GCN_LOSS_WEIGHT * F.cross_entropy(logits, (labels - shift).long())
        self.last["nodes_out"] = out
        return node_loss, proto

    # ---- paradigm update (a8; condgraph.py:558-617) ------------------
    @torch.no_grad()
    def update_prototype(self, batch, momentum=0.95):
        mh = self.mh
        batch = batch.detach()
        exist = batch.sum(-1).bool()
        if not bool(exist.any()):
            if mh.USE_RNN:
                it = self.counter_rnn()
                if it == self.P:
                    for i in range(it - 1):
                        self.prototype[:, :, i] = self.prototype[:, :, i + 1].clone()
            elif self.P > 1:
                self.counter()
            return
        if self.P == 1:
            old = self.prototype[exist]
            m = F.cosine_similarity(old, batch[exist]).unsqueeze(1) if mh.COSINE_UPDATE_ON else momentum
            self.prototype[exist] = old * m + batch[exist] * (1 - m)
            return
        if mh.USE_RNN:
            it = self.counter_rnn()
            slot = it - 1 if it == self.P else it
            m = F.cosine_similarity(self.prototype[exist, :, slot], batch[exist]).unsqueeze(1) \
                if mh.COSINE_UPDATE_ON else momentum
            if it == self.P:
                for i in range(it - 1):
                    self.prototype[:, :, i] = self.prototype[:, :, i + 1].clone()
            self.prototype[exist, :, slot] = self.prototype[exist, :, slot] * m + batch[exist] * (1 - m)
            return
        slot = self.counter()
        m = F.cosine_similarity(self.prototype[exist, :, slot], batch[exist]).unsqueeze(1) \
            if mh.COSINE_UPDATE_ON else momentum
        self.prototype[exist, :, slot] = self.prototype[exist, :, slot] * m + batch[exist] * (1 - m)

    # ---- act loss (a11; condgraph.py:338-370) -----------------------------------------
    def act_loss(self, logits_per_level, labels_per_level):
        mh = self.mh
        flat = torch.cat([lg.permute(0, 2, 3, 1).reshape(-1, self.K) for lg in logits_per_level])
        lab = torch.cat([l.reshape(-1) for l in labels_per_level]).long()
        if mh.ACT_LOSS == "softmaxFL":
            return mh.ACT_LOSS_WEIGHT * softmax_focal_loss(flat, lab)
        if mh.ACT_LOSS == "sigmoidFL":
            onehot = torch.zeros(lab.numel(), 2, device=flat.device)
            onehot[torch.arange(lab.numel(), device=flat.device), lab] = 1
            return mh.ACT_LOSS_WEIGHT * bce_focal_loss(flat, onehot)
        return None

    def post(self, feats, acts):
        if self.mh.CAT_ACT_MAP:
            return self.head_out([torch.cat([f, a], 1) for f, a in zip(feats, acts)])
        return feats

    # ---- transfer losses (a14; condgraph.py:457-498) -----------------
    def transfer_loss(self, tg_proto, tg_nodes, tg_labels):
        cfgt = self.mh.TRANSFER_CFG
        sr = (self.prototype.mean(-1) if self.P > 1 else self.prototype).detach()
        total = None

        def add(v):
            nonlocal total
            total = v if total is None else total + v

        if "NODES" in cfgt or "NODE" in cfgt:
            logp = tg_nodes.softmax(-1).log()
            tgt = sr[tg_labels.long()].softmax(-1)
            add((torch.xlogy(tgt, tgt) - tgt * logp).mean())           # nn.KLDivLoss(reduction='mean')
        if "PROTOTYPE" in cfgt:
            idx = tg_proto.sum(-1).bool()
            logp = tg_proto[idx].softmax(-1).log()
            tgt = sr[idx].softmax(-1)
            add((torch.xlogy(tgt, tgt) - tgt * logp).mean())
        if "ADJ" in cfgt:
            idx = tg_proto.sum(-1).bool()
            a = cosine_matrix(sr[idx], sr[idx]).reshape(1, -1)
            b = cosine_matrix(tg_proto[idx], tg_proto[idx]).reshape(1, -1)
            add((1 - F.cosine_similarity(a, b, dim=1, eps=1e-8)).mean())   # CosineEmbeddingLoss(margin 0), y=1
        if "ADJ_COMPLETE" in cfgt:
            idx = ~tg_proto.sum(-1).bool()
            comp = tg_proto
            comp[idx] = sr[idx]                                        # in-place, as the reference
            a = cosine_matrix(sr, sr).reshape(1, -1)
            b = cosine_matrix(comp, comp).reshape(1, -1)
            add((1 - F.cosine_similarity(a, b, dim=1, eps=1e-8)).mean())
        return total

    # ---- branches ---------------------------------------------------------------------
    def forward(self, images, features, targets=None, return_maps=False, mode="source", forward_target=False):
        feats = self.head_in(list(features))
        self.last = {"features_in": feats}
        mh = self.mh
        if self.training and targets and mode == "source":
            shapes = [tuple(f.shape[-2:]) for f in feats]
            boxes = [t.bbox for t in targets]
            labs = [t.get_field("labels") for t in targets]
            labels = fcos_assign(shapes, self.strides, boxes, labs)
            lv, rows, nlab = source_node_indices(labels, self.with_bg)
            nodes = gather_nodes(feats, lv, rows)
            self.last.update(labels=labels, node_level=lv, node_rows=rows, node_labels=nlab, nodes=nodes)
            node_loss, proto = self.forward_gcns(nodes, nlab)
            self.last["prototype_batch"] = proto
            self.update_prototype(proto)
            w = self.conded_weight()
            self.last["conded_weight"] = w
            logits = [self.dynamic_conv(f, w) for f in feats]
            acts = [self.act_of(lg) for lg in logits]
            loss = self.act_loss(logits, labels) if mh.ACT_LOSS else None
            return self.post(feats, acts), (node_loss, 0), loss, acts
        if self.training and mode == "target" and forward_target:
            acts = [self.act_of(self.dynamic_conv(f, self.conded_weight())) for f in feats]
            lv, rows, plab, masks = target_node_indices(
                acts, feats, mh.DBSCAN_EPS, mh.DBSCAN_THR, self.use_sklearn,
                mh.TARGET_SAMPLING_CFG, self.cfg.SOLVER.MIDDLE_HEAD.PLABEL_TH[0])
            self.last.update(node_level=lv, node_rows=rows, node_labels=plab)
            if mh.TARGET_SAMPLING_CFG == "dbscan":
                self.last["dbscan_masks"] = masks
            out = self.post(feats, acts)
            if lv is not None and (mh.TRANSFER_CFG[0] is not None or mh.GCN_SELF_TRAINING):
                nodes = gather_nodes(feats, lv, rows)
                node_loss, tg_proto = self.forward_gcns(nodes, plab)
                node_loss = mh.GCN_LOSS_WEIGHT_TG * node_loss
                # with GLOBAL_GCN=False the reference has overwritten the sampled rows IN PLACE with the
                # GCN outputs (condgraph.py:413) before it reaches get_transfer_loss (:526)
                tl_nodes = nodes if mh.GLOBAL_GCN else self.last["nodes_out"]
                tl = self.transfer_loss(tg_proto, tl_nodes, plab)
                if tl is not None and bool(tl):
                    tl = mh.CON_LOSS_WEIGHT * tl
                if mh.GCN_SELF_TRAINING:
                    return out, (node_loss, tl), None, acts
                return out, (None, tl), None, acts
            return out, None, None, acts
        w = self.conded_weight()
        acts = [self.act_of(self.dynamic_conv(f, w)) for f in feats]
        return self.post(feats, acts), None, None, acts


def build_oracle(cfg, in_channels=256):
    return OracleCondGraph(cfg, in_channels)


# ----------------------------------------------------------------------------------------------------------------------
# f3: CKA discriminator (modeling/discriminator/fcos_head_discriminator_con.py:88-127, layer.py:6-24), restated with plain
# torch ops.  TEST INFRASTRUCTURE (pinned by tests/golden/cka.npz, generated from the unmodified reference module).
# ----------------------------------------------------------------------------------------------------------------------
def cka_discriminator_loss(state, feature, act_maps, target, num_classes_fg, num_convs, fusion_cfg="concat"):
    """state: the module's state dict (dis_tower.{3i}.weight/bias, dis_tower.{3i+1}.weight/bias, classifier_cls_{c}.{0,2}.*).
    Forward only (the gradient reversal is the identity forward, layer.py:15-17); gradients come from autograd on the inputs,
    with the reversal applied by the caller (d = -lambda * d, layer.py:19-24)."""
    x = feature
    for i in range(num_convs):                                                      # :20-33, :98
        x = F.conv2d(x, state["dis_tower.%d.weight" % (3 * i)], state["dis_tower.%d.bias" % (3 * i)], padding=1)
        x = F.relu(F.group_norm(x, 32, state["dis_tower.%d.weight" % (3 * i + 1)], state["dis_tower.%d.bias" % (3 * i + 1)], 1e-5))
    loss = 0
    for c in range(num_classes_fg):                                                 # :101-123 (use_bg = False: map index c + 1)
        m = act_maps[:, c + 1:c + 2]
        if fusion_cfg != "concat":
            raise KeyError("only 'concat' is restated")
        x_cls = torch.cat((x, m), dim=1)                                            # :104-105
        hcls = F.relu(F.conv2d(x_cls, state["classifier_cls_%d.0.weight" % c], state["classifier_cls_%d.0.bias" % c], padding=1))
        logits = F.conv2d(hcls, state["classifier_cls_%d.2.weight" % c], state["classifier_cls_%d.2.bias" % c], padding=1)
        targets = torch.full(logits.shape, float(target), dtype=torch.float, device=x.device)
        if num_classes_fg > 1:                                                      # :115-118
            loss_cls = F.binary_cross_entropy_with_logits(logits, targets, weight=m.detach(), reduction="sum") / m.sum().detach()
        else:                                                                       # :119-120
            loss_cls = F.binary_cross_entropy_with_logits(logits, targets)
        loss = loss + loss_cls / num_classes_fg                                     # :121
    return loss
