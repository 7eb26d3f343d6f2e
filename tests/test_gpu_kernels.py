"""Kernel-level parity (B200): each C-ABI entry point against the CPU oracle / plain torch fp32 on seeded inputs.
Bit-exact for labels, indices and DBSCAN labels; rtol 1e-3 for the tensor-core path, tighter for fp32 kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from scan_b200 import ops  # noqa: E402
from scan_b200.synthetic import make_boxes, make_features  # noqa: E402
from oracle import condgraph_oracle as orc  # noqa: E402

SHAPES = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)]
FULL = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
STRIDES = [8, 16, 32, 64, 128]
DEV = "cuda"


def _close(a, b, rtol, what, atol=5e-6):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = max(float(b.abs().max()), 1e-30)
    err = float((a - b).abs().max())
    assert err <= rtol * scale + atol, "%s: max|d|=%.3e scale=%.3e" % (what, err, scale)


@pytest.mark.parametrize("shapes,n", [(SHAPES, 3), (FULL, 2)])
def test_pack_unpack_rows(shapes, n):
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(n, 256, h, w, generator=g) for h, w in shapes]
    geo = ops.Geometry(shapes, STRIDES, n)
    dfeats = [f.to(DEV).requires_grad_(True) for f in feats]
    rows = ops.pack_rows(geo, dfeats)
    want = torch.cat([orc.nhwc_rows(f) for f in feats])
    assert torch.equal(rows.cpu(), want)
    cot = torch.randn(rows.shape, generator=g)
    rows.backward(cot.to(DEV))
    for l, f in enumerate(dfeats):
        a, b = geo.row_off[l], geo.row_off[l + 1]
        h, w = shapes[l]
        assert torch.equal(f.grad.cpu(), cot[a:b].reshape(n, h, w, 256).permute(0, 3, 1, 2))


@pytest.mark.parametrize("shapes,hw,n,nb,seed", [(SHAPES, (200, 336), 3, 6, 1), (FULL, (800, 1344), 2, 18, 2),
                                                 (FULL, (800, 1344), 8, 18, 1234), (SHAPES, (200, 336), 2, 1, 5)])
def test_fcos_assign_and_source_sampling_bit_exact(shapes, hw, n, nb, seed):
    boxes = make_boxes(n, 8, nb, hw, seed)
    # ragged box counts + a duplicated box (area tie -> first index wins)
    bl = [(b[: max(1, nb - i)], l[: max(1, nb - i)]) for i, (b, l) in enumerate(boxes)]
    if nb > 2:
        b0, l0 = bl[0]
        bl[0] = (torch.cat([b0, b0[:1]]), torch.cat([l0, (l0[:1] % 8) + 1]))
    geo = ops.Geometry(shapes, STRIDES, n)

    class T(object):
        mode = "xyxy"

        def __init__(self, b, l):
            self.bbox, self._l = b, l

        def get_field(self, k):
            return self._l

    pb, pl, pc, gmax = ops.pad_targets([T(b, l) for b, l in bl], DEV)
    labels = ops.fcos_assign(geo, pb, pl, pc, gmax)
    want = orc.fcos_assign(shapes, STRIDES, [b for b, _ in bl], [l for _, l in bl])
    assert torch.equal(labels.cpu(), torch.cat(want))
    for with_bg in (True, False):
        smp = ops.sample_nodes(geo, 0, with_bg, labels=labels)
        lv, rows, labs = orc.source_node_indices(want, with_bg)
        off = torch.tensor(geo.row_off[:-1])
        assert torch.equal(smp.node_rows.cpu().long(), rows + off[lv])
        assert torch.equal(smp.node_labels.cpu(), labs)


def test_target_sampling_edge_cases_bit_exact():
    shapes = [(6, 7), (3, 4), (2, 2)]
    n = 2
    geo = ops.Geometry(shapes, STRIDES[:3], n)
    rs = np.random.RandomState(3)
    for trial in range(12):
        masks = []
        for l, (h, w) in enumerate(shapes):
            m = rs.rand(n * h * w) < [0.3, 0.9, 0.0][(l + trial) % 3]
            if trial == 5 and l == 0:
                m[:] = True
                m[7] = False          # n_neg == 1: negative indices wrap
            masks.append(m)
        mask = torch.from_numpy(np.concatenate(masks).astype(np.uint8))
        plabel = torch.from_numpy(rs.randint(1, 9, geo.R).astype(np.int64))
        if any(m.all() for m in masks):   # a level without any negative: the reference raises (loss.py:503-504)
            with pytest.raises(IndexError):
                ops.sample_nodes(geo, 1, True, pos_mask=mask.to(DEV), plabel=plabel.to(DEV))
            continue
        smp = ops.sample_nodes(geo, 1, True, pos_mask=mask.to(DEV), plabel=plabel.to(DEV))
        rows_w, labs_w = [], []
        neg_w, negl_w = [], []
        for l in range(len(shapes)):
            a = geo.row_off[l]
            mk = masks[l]
            p = np.nonzero(mk)[0]
            q = np.nonzero(~mk)[0]
            if len(p) == 0:
                continue
            k = orc.floor_linspace(len(q) - 2, len(p))
            neg_w.append(q[k] + a)
            negl_w.append(np.zeros(len(p), np.int64))
            rows_w.append(p + a)
            labs_w.append(plabel.numpy()[p + a])
        if not rows_w:
            assert smp.n_nodes == 0
            continue
        assert np.array_equal(smp.node_rows.cpu().numpy(), np.concatenate(neg_w + rows_w))
        assert np.array_equal(smp.node_labels.cpu().numpy(), np.concatenate(negl_w + labs_w))
    # a level whose locations are all positive: the reference raises IndexError (loss.py:503-504)
    mask = torch.ones(geo.R, dtype=torch.uint8)
    with pytest.raises(IndexError):
        ops.sample_nodes(geo, 1, True, pos_mask=mask.to(DEV), plabel=torch.ones(geo.R, dtype=torch.int64, device=DEV))


def test_gather_scatter_rows():
    g = torch.Generator().manual_seed(1)
    rows = torch.randn(5000, 256, generator=g)
    idx = torch.randint(0, 5000, (1777,), generator=g).sort().values
    r = rows.to(DEV).requires_grad_(True)
    out = ops.gather_rows(r, idx.to(DEV).int())
    assert torch.equal(out.cpu(), rows[idx])
    cot = torch.randn(out.shape, generator=g)
    out.backward(cot.to(DEV))
    want = torch.zeros_like(rows).index_add_(0, idx, cot)
    _close(r.grad, want, 1e-6, "scatter_add")


@pytest.mark.parametrize("shapes,n", [(SHAPES, 3), (FULL, 2), ([(1, 1), (5, 3)], 1)])
def test_gn_relu_levels_matches_torch_group_norm(shapes, n):
    """f1: GroupNorm(32)+ReLU over all levels in the rows layout vs torch's group_norm + relu, forward and backward."""
    g = torch.Generator().manual_seed(11)
    geo = ops.Geometry(shapes, STRIDES[:len(shapes)], n)
    xs = [(torch.randn(n, 256, h, w, generator=g) * 2 + 0.3) for h, w in shapes]
    gamma, beta = torch.randn(256, generator=g), torch.randn(256, generator=g) * 0.1
    cots = [torch.randn(n, 256, h, w, generator=g) for h, w in shapes]
    xr = [x.clone().requires_grad_(True) for x in xs]
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    cb = torch.randn(256, generator=g) * 0.5            # bias of the preceding convolution, folded into the kernel
    cbr = cb.clone().requires_grad_(True)
    yr = [torch.relu(torch.nn.functional.group_norm(x + cbr.view(1, -1, 1, 1), 32, gr, br, 1e-5)) for x in xr]
    sum((y * c).sum() for y, c in zip(yr, cots)).backward()
    xd = [x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for x in xs]
    gd, bd = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    cbd = cb.to(DEV).requires_grad_(True)
    yd = ops.gn_relu_levels(geo, gd, bd, 1e-5, xd, conv_bias=cbd)
    rows = ops.join_rows(geo, yd)                       # zero-copy: the levels are adjacent views of one buffer
    assert rows.data_ptr() == yd[0].data_ptr() and rows.shape == (geo.R, 256)
    (sum((y * c.to(DEV)).sum() for y, c in zip(yd, cots)) + 0.0 * rows.sum()).backward()
    for l in range(len(shapes)):
        _close(yd[l], yr[l], 2e-5, "y_l%d" % l)
        _close(xd[l].grad, xr[l].grad, 5e-5, "dx_l%d" % l)
        want_rows = yr[l].permute(0, 2, 3, 1).reshape(-1, 256)
        assert torch.allclose(rows[geo.row_off[l]:geo.row_off[l + 1]].cpu(), want_rows.detach(), rtol=1e-4, atol=1e-5)
    _close(gd.grad, gr.grad, 5e-5, "dgamma")
    _close(bd.grad, br.grad, 5e-5, "dbeta")
    _close(cbd.grad, cbr.grad, 5e-5, "d_conv_bias", atol=2e-4)      # sums to ~0 analytically within each group


@pytest.mark.parametrize("shapes,n", [(SHAPES, 2), ([(1, 1), (5, 3)], 1)])
def test_add_relu_levels(shapes, n):
    g = torch.Generator().manual_seed(13)
    geo = ops.Geometry(shapes, STRIDES[:len(shapes)], n)
    us = [torch.randn(n, 256, h, w, generator=g) for h, w in shapes]
    vs = [torch.randn(n, 256, h, w, generator=g) for h, w in shapes]
    bias = torch.randn(256, generator=g)
    cots = [torch.randn(n, 256, h, w, generator=g) for h, w in shapes]
    ur, vr, br = [t.clone().requires_grad_(True) for t in us], [t.clone().requires_grad_(True) for t in vs], bias.clone().requires_grad_(True)
    yr = [torch.relu(u + v + br.view(1, -1, 1, 1)) for u, v in zip(ur, vr)]
    sum((y * c).sum() for y, c in zip(yr, cots)).backward()
    ud = [t.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for t in us]
    vd = [t.to(DEV).requires_grad_(True) for t in vs]            # NCHW-contiguous inputs are converted inside
    bd = bias.to(DEV).requires_grad_(True)
    yd = ops.add_relu_levels(geo, bd, ud, vd)
    sum((y * c.to(DEV)).sum() for y, c in zip(yd, cots)).backward()
    for l in range(len(shapes)):
        _close(yd[l], yr[l], 1e-6, "y_l%d" % l)            # (u + bias) + v here, (u + v) + bias in torch: last-bit differences
        flip = (yr[l].detach().abs() < 1e-5)                # the ReLU mask may differ where the pre-activation is ~0
        assert torch.equal(ud[l].grad.cpu()[~flip], ur[l].grad[~flip]) and torch.equal(vd[l].grad.cpu()[~flip], vr[l].grad[~flip])
    _close(bd.grad, br.grad, 1e-4, "d_bias")


def test_pack_levels_roundtrip_and_gradient():
    g = torch.Generator().manual_seed(12)
    geo = ops.Geometry(SHAPES, STRIDES, 2)
    feats = [torch.randn(2, 256, h, w, generator=g) for h, w in SHAPES]
    fd = [f.to(DEV).requires_grad_(True) for f in feats]
    levels = ops.pack_levels(geo, fd)
    for l, f in enumerate(feats):
        assert levels[l].shape == f.shape and levels[l].permute(0, 2, 3, 1).is_contiguous()
        assert torch.equal(levels[l].cpu(), f)
    rows = ops.join_rows(geo, levels)
    assert rows.data_ptr() == levels[0].data_ptr()
    cot = torch.randn(rows.shape, generator=g)
    (rows * cot.to(DEV)).sum().backward()
    for l, f in enumerate(fd):
        h, w = SHAPES[l]
        want = cot[geo.row_off[l]:geo.row_off[l + 1]].view(2, h, w, 256).permute(0, 3, 1, 2)
        assert torch.equal(f.grad.cpu(), want)


def test_join_rows_copies_when_levels_are_not_adjacent():
    """join_rows is zero-copy only for adjacent views of one buffer; anything else (torch-produced tower outputs of the IN / BN
    tower variants, NCHW-contiguous tensors) goes through one concatenating copy with the same values and gradients."""
    g = torch.Generator().manual_seed(14)
    geo = ops.Geometry(SHAPES, STRIDES, 2)
    feats = [torch.randn(2, 256, h, w, generator=g) for h, w in SHAPES]
    fd = [f.to(DEV).requires_grad_(True) for f in feats]                 # separately allocated, NCHW-contiguous
    rows = ops.join_rows(geo, fd)
    want = torch.cat([f.permute(0, 2, 3, 1).reshape(-1, 256) for f in feats])
    assert torch.equal(rows.cpu(), want)
    cot = torch.randn(rows.shape, generator=g)
    (rows * cot.to(DEV)).sum().backward()
    for l, f in enumerate(fd):
        h, w = SHAPES[l]
        assert torch.equal(f.grad.cpu(), cot[geo.row_off[l]:geo.row_off[l + 1]].view(2, h, w, 256).permute(0, 3, 1, 2))


def _condconv_reference(feats, weight, bias, labels, mode, lam):
    k = weight.shape[0]
    logits = [torch.nn.functional.conv2d(f, weight.reshape(k, -1, 1, 1), bias) for f in feats]
    acts = [lg.softmax(1) if mode == 0 else lg.sigmoid() for lg in logits]
    loss = None
    if labels is not None:
        flat = torch.cat([lg.permute(0, 2, 3, 1).reshape(-1, k) for lg in logits])
        if mode == 0:
            loss = lam * orc.softmax_focal_loss(flat, labels)
        else:
            onehot = torch.zeros(labels.numel(), 2, dtype=flat.dtype)
            onehot[torch.arange(labels.numel()), labels] = 1
            loss = lam * orc.bce_focal_loss(flat, onehot)
    return acts, loss


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("k,mode,with_bias,shapes,n", [(9, 0, False, SHAPES, 2), (2, 1, False, SHAPES, 3), (9, 0, True, SHAPES, 1),
                                                       (2, 0, False, SHAPES, 2), (9, 0, False, FULL, 1), (16, 0, False, SHAPES, 1)])
def test_condconv_forward_backward(impl, k, mode, with_bias, shapes, n):
    """tcgen05 tf32 path (impl 0) within rtol 1e-3 of the fp32 torch reference; FFMA kernel (impl 1) within 1e-5."""
    rtol = 2e-5 if impl == 1 else 1e-3   # impl 0: tcgen05 3xTF32 (product), impl 1: fp32 FFMA verification kernel
    g = torch.Generator().manual_seed(10 + k)
    feats = [torch.relu(torch.randn(n, 256, h, w, generator=g)) for h, w in shapes]
    weight = torch.randn(k, 256, generator=g) * 0.08
    bias = torch.randn(k, generator=g) if with_bias else None
    geo = ops.Geometry(shapes, STRIDES, n)
    labels = torch.randint(0, k, (geo.R,), generator=g)
    lam = 0.7
    # reference (CPU fp64-free fp32 torch)
    fr = [f.clone().requires_grad_(True) for f in feats]
    wr = weight.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True) if with_bias else None
    acts_r, loss_r = _condconv_reference(fr, wr, br, labels, mode, lam)
    cots = [torch.randn(a.shape, generator=g) / a.numel() * 50 for a in acts_r]
    (loss_r * 1.3 + sum((a * c).sum() for a, c in zip(acts_r, cots))).backward()
    # device
    old = ops.CONDCONV_IMPL["impl"]
    ops.CONDCONV_IMPL["impl"] = impl
    try:
        fd = [f.to(DEV).requires_grad_(True) for f in feats]
        wd = weight.to(DEV).requires_grad_(True)
        bd = bias.to(DEV).requires_grad_(True) if with_bias else None
        rows = ops.pack_rows(geo, fd)
        acts_d, loss_d, _ = ops.condconv(geo, rows, wd, bd, k, mode, labels.to(DEV), lam)
        (loss_d * 1.3 + sum((a * c.to(DEV)).sum() for a, c in zip(acts_d, cots))).backward()
    finally:
        ops.CONDCONV_IMPL["impl"] = old
    for l in range(len(shapes)):
        _close(acts_d[l], acts_r[l], rtol, "act level %d" % l)
        _close(fd[l].grad, fr[l].grad, rtol, "d_feat level %d" % l)
    _close(loss_d, loss_r, rtol, "loss")
    _close(wd.grad, wr.grad, rtol, "d_weight")
    if with_bias:
        _close(bd.grad, br.grad, rtol, "d_bias")


@pytest.mark.parametrize("m,drop", [(1, 0.0), (10, 0.0), (13, 0.0), (64, 0.0), (257, 0.0), (1048, 0.0)])
def test_attention_matches_reference_view_semantics(m, drop):
    g = torch.Generator().manual_seed(m)
    q, k, v = [torch.randn(m, 256, generator=g) for _ in range(3)]
    qr, kr, vr = [t.clone().requires_grad_(True) for t in (q, k, v)]
    att = torch.softmax(torch.bmm(qr.reshape(4, m, 64), kr.reshape(4, m, 64).transpose(1, 2)) * 0.25, dim=2)
    ctx_r = torch.bmm(att, vr.reshape(4, m, 64)).reshape(m, 256)
    cot = torch.randn(m, 256, generator=g)
    (ctx_r * cot).sum().backward()
    qd, kd, vd = [t.to(DEV).requires_grad_(True) for t in (q, k, v)]
    ctx_d = ops.chunked_attention(qd, kd, vd, 0.25)
    (ctx_d * cot.to(DEV)).sum().backward()
    _close(ctx_d, ctx_r, 2e-5, "ctx")
    # m == 1: dq, dk are exactly zero (softmax over one key); what is left is fp32 rounding of dP - D ~ 1e-6 |dO||V|
    _close(qd.grad, qr.grad, 5e-5, "dq", atol=2e-5)
    _close(kd.grad, kr.grad, 5e-5, "dk", atol=2e-5)
    _close(vd.grad, vr.grad, 5e-5, "dv")


def test_attention_dropout_is_consistent_between_forward_and_backward():
    m = 300
    g = torch.Generator().manual_seed(0)
    q, k, v = [torch.randn(m, 256, generator=g).to(DEV) for _ in range(3)]
    v1 = v.clone().requires_grad_(True)
    out = ops.chunked_attention(q, k, v1, 0.25, 0.1, 1234)
    out2 = ops.chunked_attention(q, k, v, 0.25, 0.1, 1234)
    assert torch.equal(out, out2)                       # same seed -> same mask
    out3 = ops.chunked_attention(q, k, v, 0.25, 0.1, 99)
    assert not torch.equal(out, out3)
    # out is linear in v for a fixed mask: d(sum(out*c))/dv evaluated by finite differences of the kernel itself
    c = torch.randn(m, 256, generator=g).to(DEV)
    (out * c).sum().backward()
    dv = torch.zeros_like(v)
    dv[5, 7] = 1.0
    fd = ((ops.chunked_attention(q, k, v + dv, 0.25, 0.1, 1234) - out2) * c).sum()
    _close(v1.grad[5, 7], fd, 2e-3, "dropout dv")
    # mean preserved: E[mask/keep] = 1
    base = ops.chunked_attention(q, k, torch.ones_like(v), 0.25, 0.0, 0)
    dropped = ops.chunked_attention(q, k, torch.ones_like(v), 0.25, 0.1, 7)
    assert abs(float(dropped.mean()) - float(base.mean())) < 0.02


@pytest.mark.parametrize("m,drop", [(300, 0.1), (1, 0.0), (129, 0.3), (1000, 0.0)])
def test_attention_tcgen05_and_ffma_kernels_agree(m, drop):
    """The tcgen05 forward/backward (product) against the fp32 FFMA kernels, same dropout mask (shared counter hash)."""
    g = torch.Generator().manual_seed(m + 17)
    q, k, v, cot = [torch.randn(m, 256, generator=g).to(DEV) for _ in range(4)]
    res = {}
    saved = dict(ops.ATTN_IMPL)
    try:
        for impl in ("ffma", "t5"):
            ops.ATTN_IMPL.update(fwd=impl, bwd=impl)
            qd, kd, vd = [t.clone().requires_grad_(True) for t in (q, k, v)]
            out = ops.chunked_attention(qd, kd, vd, 0.25, drop, 4321)
            (out * cot).sum().backward()
            res[impl] = (out.detach(), qd.grad, kd.grad, vd.grad)
    finally:
        ops.ATTN_IMPL.update(saved)
    for name, a, b in zip(("ctx", "dq", "dk", "dv"), res["t5"], res["ffma"]):
        _close(a, b, 5e-5, name, atol=2e-5)


@pytest.mark.parametrize("k,p", [(9, 3), (2, 1), (16, 5)])
def test_manifest_rnn_matches_torch(k, p):
    """K4a: nn.RNN(256, 512, 2, tanh) over the paradigm slots + the (P x 1) convolution, forward and all parameter gradients."""
    torch.manual_seed(3 + k)
    rnn = torch.nn.RNN(256, 512, 2, nonlinearity="tanh")
    conv = torch.nn.Conv2d(512, 256, kernel_size=(p, 1))
    proto = torch.randn(k, 256, p)
    cot = torch.randn(k, 256)
    h, _ = rnn(proto.permute(2, 0, 1).contiguous())
    want = torch.einsum("pkc,ocp->ko", h, conv.weight[:, :, :, 0]) + conv.bias
    (want * cot).sum().backward()
    ref = {n: t.grad.clone() for n, t in list(rnn.named_parameters()) + [("wc", conv.weight), ("bc", conv.bias)]}
    rnn_d, conv_d = torch.nn.RNN(256, 512, 2, nonlinearity="tanh").to(DEV), torch.nn.Conv2d(512, 256, kernel_size=(p, 1)).to(DEV)
    rnn_d.load_state_dict(rnn.state_dict())
    conv_d.load_state_dict(conv.state_dict())
    got = ops.manifest_rnn(proto.to(DEV), rnn_d, conv_d)
    (got * cot.to(DEV)).sum().backward()
    _close(got, want, 2e-5, "kernel")
    for n, t in list(rnn_d.named_parameters()) + [("wc", conv_d.weight), ("bc", conv_d.bias)]:
        _close(t.grad, ref[n], 5e-5, "d_" + n)


def test_class_sums_and_proto_update():
    g = torch.Generator().manual_seed(4)
    m, k = 3000, 9
    nodes = torch.randn(m, 256, generator=g)
    labels = torch.randint(0, k, (m,), generator=g)
    labels[labels == 3] = 4          # class 3 absent
    packed = ops.class_sums(nodes.to(DEV), labels.to(DEV), k, 0).cpu()
    for c in range(k):
        sel = nodes[labels == c]
        assert float(packed[c, 256]) == sel.shape[0]
        _close(packed[c, :256], sel.sum(0) if sel.shape[0] else torch.zeros(256), 1e-5 if sel.shape[0] else 1, "class sum")
    for p, slot, shift, cosine in [(3, 0, False, True), (3, 2, True, True), (3, 1, False, False), (1, 0, False, True)]:
        proto = torch.randn(k, 256, p, generator=g) if p > 1 else torch.randn(k, 256, generator=g)
        want = proto.clone()
        batch = torch.zeros(k, 256)
        for c in range(k):
            if (labels == c).any():
                batch[c] = nodes[labels == c].mean(0)
        exist = batch.sum(-1).bool()
        old = want[exist, :, slot] if p > 1 else want[exist]
        mom = torch.cosine_similarity(old, batch[exist]).unsqueeze(1) if cosine else 0.95
        if shift:
            for i in range(p - 1):
                want[:, :, i] = want[:, :, i + 1].clone()
        new = old * mom + batch[exist] * (1 - mom)
        if p > 1:
            want[exist, :, slot] = new
        else:
            want[exist] = new
        dproto = proto.to(DEV).contiguous()
        got_batch = ops.proto_update(ops.class_sums(nodes.to(DEV), labels.to(DEV), k, 0), dproto, slot, shift, cosine, 0.95)
        _close(got_batch, batch, 1e-5, "prototype_batch")
        _close(dproto, want, 1e-5, "prototype p=%d slot=%d" % (p, slot))


def _blobs(rs, n, kind):
    if kind == "planar":
        p2 = rs.uniform(0, 0.12 * np.sqrt(n) * 3, (n, 2))
        basis = np.linalg.qr(rs.standard_normal((256, 2)))[0]
        return (p2 @ basis.T).astype(np.float32)
    if kind == "dup":
        base = rs.standard_normal((max(n // 7, 1), 256)).astype(np.float32) * 0.2
        return base[rs.randint(0, base.shape[0], n)]
    centers = rs.standard_normal((5, 256)) * 3
    return (centers[rs.randint(0, 5, n)] + rs.standard_normal((n, 256)) * 0.17).astype(np.float32)


@pytest.mark.parametrize("n,kind,eps", [(1, "blob", 3.0), (4, "blob", 3.0), (63, "blob", 3.0), (64, "planar", 3.0), (65, "dup", 0.5),
                                        (500, "planar", 3.0), (777, "blob", 3.0), (2000, "planar", 2.0), (3000, "blob", 4.0),
                                        (1500, "planar", 0.7)])
def test_dbscan_points_bit_exact_vs_sklearn(n, kind, eps):
    from sklearn.cluster import DBSCAN
    rs = np.random.RandomState(n)
    x = _blobs(rs, n, kind)
    want = DBSCAN(eps=eps).fit_predict(x)
    labels, info = ops.dbscan_points(torch.from_numpy(x).to(DEV), eps)
    got = labels.cpu().numpy()
    assert np.array_equal(got, want), "mismatch at %s" % np.nonzero(got != want)[0][:10]
    info = info.cpu().numpy()
    assert info[1] == want.max() + 1 and info[2] == int((want < 0).sum())


def test_dbscan_threshold_band_is_rechecked_in_fp64():
    """Points at distance exactly eps (and one ulp either side) from a dense core: fp32 alone would flip them."""
    from sklearn.cluster import DBSCAN
    rs = np.random.RandomState(0)
    core = np.zeros((8, 256), np.float32)
    core[:, 0] = rs.uniform(0, 1e-3, 8)
    ring = np.zeros((40, 256), np.float32)
    for i in range(40):
        d = rs.standard_normal(255)
        d = d / np.linalg.norm(d) * 3.0
        ring[i, 1:] = d.astype(np.float32) * np.float32(1 + (i - 20) * 1e-7)
    x = np.concatenate([core, ring])
    want = DBSCAN(eps=3.0).fit_predict(x)
    labels, info = ops.dbscan_points(torch.from_numpy(x).to(DEV), 3.0)
    assert np.array_equal(labels.cpu().numpy(), want)
    assert int(info.cpu()[5]) > 0   # the fp64 path was actually taken


@pytest.mark.parametrize("r,c", [(5000, 8), (179200, 8), (4097, 2), (333, 1), (1000, 5), (64, 4)])
def test_sigmoid_focal_rows_kernels(r, c):
    """a17: thread-per-row kernels (C = 8 / 4 / 2 / 1 vectorised, generic C) vs the reference formulas, incl. ignored rows (t < 0),
    background rows (t = 0) and out-of-range class ids."""
    g = torch.Generator().manual_seed(2 + c)
    logits = torch.randn(r, c, generator=g) * 3
    logits[0, 0], logits[1, 0] = 95.0, -95.0           # p saturates / underflows: the reference's log(max(p, FLT_MIN)) clamp
    targets = torch.randint(-1, c + 2, (r,), generator=g).int()
    want = orc.sigmoid_focal_loss_elementwise(logits, targets, 2.0, 0.25)
    cot = torch.randn(r, c, generator=g)
    # the reference's backward is its own analytic kernel (SigmoidFocalLoss_cuda.cu:61-101), not autograd of the forward
    want_grad = orc.sigmoid_focal_loss_backward_elementwise(logits, targets, cot, 2.0, 0.25)
    ld = logits.to(DEV).requires_grad_(True)
    got = ops.sigmoid_focal_loss(ld, targets.to(DEV), 2.0, 0.25)
    (got * cot.to(DEV)).sum().backward()
    _close(got, want, 1e-5, "sigmoid focal fwd")
    _close(ld.grad, want_grad, 1e-5, "sigmoid focal bwd")
    # away from the clamp the analytic backward equals autograd of the forward formula
    lr = logits[2:].clone().requires_grad_(True)
    (orc.sigmoid_focal_loss_elementwise(lr, targets[2:], 2.0, 0.25) * cot[2:]).sum().backward()
    _close(ld.grad[2:], lr.grad, 1e-5, "sigmoid focal bwd vs autograd")
    # a non-default gamma takes the powf branch
    want15 = orc.sigmoid_focal_loss_elementwise(logits[:512], targets[:512], 1.5, 0.4)
    got15 = ops.sigmoid_focal_loss(logits[:512].to(DEV), targets[:512].to(DEV), 1.5, 0.4)
    _close(got15, want15, 2e-5, "sigmoid focal gamma 1.5")


def test_ensemble_all_levels_one_launch():
    g = torch.Generator().manual_seed(2)
    acts = [torch.rand(2, 9, h, w, generator=g) for h, w in SHAPES]
    cls = [torch.randn(2, 8, h, w, generator=g) for h, w in SHAPES]
    for mode in ("common", "light", "precision"):
        want = orc.ensemble(mode, cls, acts)
        got = ops.ensemble_levels(mode, None if mode == "light" else [c.to(DEV) for c in cls], [a.to(DEV) for a in acts])
        for l in range(len(SHAPES)):
            _close(got[l], want[l], 1e-6, "ensemble %s level %d" % (mode, l))
    one = ops.ensemble("precision", cls[1].to(DEV), acts[1].to(DEV))
    _close(one, orc.ensemble("precision", [cls[1]], [acts[1]])[0], 1e-6, "single level")


# ---------------------------------------------------------------------------------------------------------------------
# K3a': tcgen05 3xTF32 dense layers (csrc/gemm.cu) against float64 torch on the same GPU
# ---------------------------------------------------------------------------------------------------------------------
class _Mha(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.dim_per_head, self.num_heads = 64, 4
        self.linear_k, self.linear_v, self.linear_q = [torch.nn.Linear(256, 256) for _ in range(3)]
        self.linear_final = torch.nn.Linear(256, 256)
        self.layer_norm = torch.nn.LayerNorm(256)


def _mha_ref64(a, x):
    m = x.shape[0]
    q, k, v = a.linear_q(x), a.linear_k(x), a.linear_v(x)
    att = torch.softmax(torch.bmm(q.reshape(4, m, 64), k.reshape(4, m, 64).transpose(1, 2)) * 0.25, dim=2)
    ctx = torch.bmm(att, v.reshape(4, m, 64)).reshape(m, 256)
    return a.layer_norm(x + a.linear_final(ctx))


@pytest.mark.parametrize("m", [1, 10, 130, 1048, 3000])
def test_graph_attention_block_matches_float64(m):
    """q/k/v projections + attention + linear_final + residual LayerNorm as ONE op (four tcgen05 launches forward) vs float64
    torch: output, d(x) and all ten parameter gradients."""
    torch.manual_seed(m)
    a = _Mha().to(DEV)
    for p_ in a.parameters():
        p_.data.mul_(1.5)
    x = torch.randn(m, 256, device=DEV)
    cot = torch.randn(m, 256, device=DEV)
    a64 = _Mha().to(DEV).double()
    a64.load_state_dict({k: v.double() for k, v in a.state_dict().items()})
    x64 = x.double().requires_grad_(True)
    y64 = _mha_ref64(a64, x64)
    (y64 * cot.double()).sum().backward()
    xd = x.clone().requires_grad_(True)
    y = ops.graph_attention(xd, a, 0.0, 0)
    (y * cot).sum().backward()
    _close(y, y64, 2e-5, "y")
    _close(xd.grad, x64.grad, 1e-4, "dx")
    for (n, p_), (_, p64) in zip(a.named_parameters(), a64.named_parameters()):
        _close(p_.grad, p64.grad, 1e-4, "d_" + n, atol=1e-5)


def test_graph_attention_dropout_backward_is_consistent():
    """Both dropouts on (attention probabilities + linear_final output): the backward must use the masks of the forward.
    Directional finite difference of the op itself (smooth for a fixed mask)."""
    torch.manual_seed(5)
    m = 400
    a = _Mha().to(DEV)
    x = torch.randn(m, 256, device=DEV)
    cot = torch.randn(m, 256, device=DEV)
    d = torch.randn(m, 256, device=DEV)
    xd = x.clone().requires_grad_(True)
    y = ops.graph_attention(xd, a, 0.1, 777)
    y2 = ops.graph_attention(x, a, 0.1, 777)
    assert torch.equal(y, y2)
    assert not torch.equal(y, ops.graph_attention(x, a, 0.1, 778))
    (y * cot).sum().backward()
    h = 1e-2
    with torch.no_grad():
        fp = (ops.graph_attention(x + h * d, a, 0.1, 777).double() * cot.double()).sum()
        fm = (ops.graph_attention(x - h * d, a, 0.1, 777).double() * cot.double()).sum()
    fd = float((fp - fm) / (2 * h))
    an = float((xd.grad.double() * d.double()).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1.0), (fd, an)


@pytest.mark.parametrize("m,k,shift", [(1, 9, 0), (77, 9, 0), (1000, 2, 0), (8888, 9, 0), (500, 8, 1)])
def test_node_classifier_loss_matches_float64(m, k, shift):
    torch.manual_seed(m + k)
    hid, out = torch.nn.Linear(256, 512).to(DEV), torch.nn.Linear(512, k).to(DEV)
    nodes = torch.randn(m, 256, device=DEV)
    labels = torch.randint(shift, k + shift, (m,), device=DEV)
    lam = 0.7
    h64, o64 = torch.nn.Linear(256, 512).to(DEV).double(), torch.nn.Linear(512, k).to(DEV).double()
    h64.load_state_dict({n: v.double() for n, v in hid.state_dict().items()})
    o64.load_state_dict({n: v.double() for n, v in out.state_dict().items()})
    n64 = nodes.double().requires_grad_(True)
    want = lam * torch.nn.functional.cross_entropy(o64(torch.relu(h64(n64))), labels - shift)
    (want * 1.7).backward()
    nd = nodes.clone().requires_grad_(True)
    got = ops.node_classifier_loss(nd, hid, out, labels, shift, lam)
    (got * 1.7).backward()
    _close(got, want, 2e-5, "loss")
    _close(nd.grad, n64.grad, 1e-4, "d_nodes", atol=1e-7)
    for name, a, b in (("w1", hid.weight, h64.weight), ("b1", hid.bias, h64.bias), ("w2", out.weight, o64.weight), ("b2", out.bias, o64.bias)):
        _close(a.grad, b.grad, 1e-4, "d_" + name, atol=1e-7)


def test_class_means_forward_and_backward():
    g = torch.Generator().manual_seed(8)
    m, k = 2000, 9
    nodes = torch.randn(m, 256, generator=g)
    labels = torch.randint(0, k, (m,), generator=g)
    labels[labels == 5] = 6          # class 5 absent
    cot = torch.randn(k, 256, generator=g)
    nr = nodes.clone().requires_grad_(True)
    want = torch.stack([nr[labels == c].mean(0) if (labels == c).any() else torch.zeros(256) for c in range(k)])
    (want * cot).sum().backward()
    nd = nodes.to(DEV).requires_grad_(True)
    means, packed = ops.class_means(nd, labels.to(DEV), k, 0)
    (means * cot.to(DEV)).sum().backward()
    _close(means, want, 1e-5, "class means")
    _close(nd.grad, nr.grad, 1e-6, "d_nodes")
    assert float(packed[5, 256]) == 0.0


def test_attention_single_pass_underflow_falls_back_to_two_pass():
    """The single-pass forward references every row to the Cauchy-Schwarz bound |q| max|k|; with huge, nearly orthogonal q and k the
    bound sits > 2^100 above the true scores, every probability underflows, attn_finish_kernel raises its flag and the two-pass
    kernel must take over (same stream, no host round trip)."""
    torch.manual_seed(3)
    m = 300
    q = torch.randn(m, 256) * 0.05
    k = torch.randn(m, 256) * 0.05
    q.view(4 * m, 64)[:, 0] += 60.0          # every 64-d sub-token: q along e0, k along e1 -> q.k ~ 0, |q||k| = 3600
    k.view(4 * m, 64)[:, 1] += 60.0
    v = torch.randn(m, 256)
    qr, kr, vr = [t.double().requires_grad_(True) for t in (q, k, v)]
    att = torch.softmax(torch.bmm(qr.reshape(4, m, 64), kr.reshape(4, m, 64).transpose(1, 2)) * 0.25, dim=2)
    want = torch.bmm(att, vr.reshape(4, m, 64)).reshape(m, 256)
    cot = torch.randn(m, 256)
    (want * cot.double()).sum().backward()
    qd, kd, vd = [t.to(DEV).requires_grad_(True) for t in (q, k, v)]
    got = ops.chunked_attention(qd, kd, vd, 0.25)
    (got * cot.to(DEV)).sum().backward()
    assert torch.isfinite(got).all()
    _close(got, want, 5e-5, "ctx (fallback)")
    _close(vd.grad, vr.grad, 5e-5, "dv (fallback)")
    _close(qd.grad, qr.grad, 2e-4, "dq (fallback)", atol=1e-4)


@pytest.mark.parametrize("cfg_names,p_iter,absent", [(("NODES", "ADJ"), 3, ()), (("NODES", "ADJ", "PROTOTYPE"), 3, (2, 5)),
                                                     (("ADJ_COMPLETE",), 1, (0, 7)), (("NODE", "PROTOTYPE", "ADJ", "ADJ_COMPLETE"), 3, (4,)),
                                                     (("PROTOTYPE",), 1, ())])
def test_transfer_losses_match_reference_formulas(cfg_names, p_iter, absent):
    """a14: scan_transfer_* vs the oracle's restatement of get_transfer_loss (condgraph.py:457-498), loss and both gradients."""
    import types
    torch.manual_seed(len(cfg_names) * 7 + p_iter)
    k, m = 9, 700
    proto = torch.randn(k, 256, p_iter) if p_iter > 1 else torch.randn(k, 256)
    nodes = torch.randn(m, 256)
    labels = torch.randint(0, k, (m,))
    tg = torch.randn(k, 256) * 0.7
    for c in absent:
        tg[c] = 0.0                  # class absent from the target batch: zero mean row (condgraph.py:395-398)
    fake = types.SimpleNamespace(mh=types.SimpleNamespace(TRANSFER_CFG=cfg_names), prototype=proto, P=p_iter)
    nr, tr = nodes.clone().requires_grad_(True), tg.clone().requires_grad_(True)
    want = orc.OracleCondGraph.transfer_loss(fake, tr * 1.0, nr, labels)      # (tr * 1.0: ADJ_COMPLETE writes in place)
    (want * 1.3).backward()
    nd, td = nodes.to(DEV).requires_grad_(True), tg.to(DEV).requires_grad_(True)
    got = ops.transfer_loss(cfg_names, nd, labels.to(DEV), td, proto.to(DEV))
    (got * 1.3).backward()
    _close(got, want, 2e-5, "transfer loss")
    if "NODES" in cfg_names or "NODE" in cfg_names:
        _close(nd.grad, nr.grad, 2e-5, "d_nodes", atol=1e-9)
    else:
        assert nd.grad is None
    if any(n in cfg_names for n in ("PROTOTYPE", "ADJ", "ADJ_COMPLETE")):
        _close(td.grad, tr.grad, 5e-5, "d_tg_proto", atol=1e-8)


@pytest.mark.parametrize("k,i,o,relu,bias", [(9, 256, 512, True, True), (9, 768, 512, False, True), (2, 512, 257, False, True),
                                             (16, 512, 256, False, False), (1, 256, 512, True, True)])
def test_rows_linear_matches_torch(k, i, o, relu, bias):
    """a9 (no-RNN manifestation): tiny-batch dense layer, forward + d_w / d_b / d_x."""
    torch.manual_seed(k * 1000 + o)
    lin = torch.nn.Linear(i, o, bias=bias)
    x = torch.randn(k, i)
    cot = torch.randn(k, o)
    xr = x.clone().requires_grad_(True)
    yr = lin(xr)
    yr = torch.relu(yr) if relu else yr
    (yr * cot).sum().backward()
    lind = torch.nn.Linear(i, o, bias=bias).to(DEV)
    lind.load_state_dict(lin.state_dict())
    xd = x.to(DEV).requires_grad_(True)
    yd = ops.rows_linear(xd, lind.weight, lind.bias if bias else None, relu)
    (yd * cot.to(DEV)).sum().backward()
    _close(yd, yr, 1e-5, "y")
    _close(xd.grad, xr.grad, 1e-5, "d_x")
    _close(lind.weight.grad, lin.weight.grad, 1e-5, "d_w")
    if bias:
        _close(lind.bias.grad, lin.bias.grad, 1e-5, "d_b")
    # the first layer of the manifestation takes the paradigm BUFFER: no d_x requested
    y2 = ops.rows_linear(x.to(DEV), lind.weight, lind.bias if bias else None, relu)
    assert torch.equal(y2, yd.detach())


@pytest.mark.parametrize("k,c", [(9, 512), (2, 512), (16, 256)])
def test_rows_gn_relu_matches_torch(k, c):
    torch.manual_seed(k + c)
    x = torch.randn(k, c) * 2 + 0.3
    gamma, beta = torch.randn(c), torch.randn(c) * 0.2
    cot = torch.randn(k, c)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.relu(torch.nn.functional.group_norm(xr, 32, gr, br, 1e-5))
    (yr * cot).sum().backward()
    xd, gd, bd = [t.to(DEV).requires_grad_(True) for t in (x, gamma, beta)]
    yd = ops.rows_gn_relu(xd, gd, bd, 32, 1e-5)
    (yd * cot.to(DEV)).sum().backward()
    _close(yd, yr, 2e-5, "y")
    _close(xd.grad, xr.grad, 5e-5, "d_x")
    _close(gd.grad, gr.grad, 5e-5, "d_gamma")
    _close(bd.grad, br.grad, 5e-5, "d_beta")


@pytest.mark.parametrize("m,k,shift,norm,act,shortcut", [(600, 9, 0, "NO", "relu", False), (1500, 9, 0, "cosine_detached", "tanh", True),
                                                         (300, 8, 1, "cosine_detached", "softmax", False), (257, 3, 0, "NO", "sigmoid", False),
                                                         (40, 9, 0, "cosine_detached", "NO", True)])
def test_local_gcn_matches_float64(m, k, shift, norm, act, shortcut):
    """a6: the per-class GCN on the scan_b200 GEMM kernels vs the oracle's restatement in float64 (forward, d(nodes), weights)."""
    torch.manual_seed(m + k)
    l1, l2 = torch.nn.Linear(256, 256), torch.nn.Linear(256, 256)
    scale = 0.05 if norm == "NO" else 1.0          # un-normalised affinities X X^T of N(0,1) rows would saturate the softmax
    nodes = torch.randn(m, 256) * scale
    labels = torch.randint(shift, k + shift, (m,))
    labels[labels == shift + 1] = shift             # one class absent
    cot = torch.randn(m, 256)
    l1d, l2d = torch.nn.Linear(256, 256).double(), torch.nn.Linear(256, 256).double()
    l1d.load_state_dict({n: v.double() for n, v in l1.state_dict().items()})
    l2d.load_state_dict({n: v.double() for n, v in l2.state_dict().items()})
    nr = nodes.double().requires_grad_(True)
    want = nr.clone()
    for c in range(k):
        idx = labels == c + shift
        if bool(idx.any()):
            sub = nr[idx]
            adj = orc.gcn_edge(sub, norm)
            want = want.index_put((torch.nonzero(idx).reshape(-1),), orc.gcn_forward(sub, adj, l1d.weight, l1d.bias, l2d.weight, l2d.bias, act, shortcut))
    (want * cot.double()).sum().backward()
    l1g, l2g = torch.nn.Linear(256, 256).to(DEV), torch.nn.Linear(256, 256).to(DEV)
    l1g.load_state_dict(l1.state_dict())
    l2g.load_state_dict(l2.state_dict())
    nd = nodes.to(DEV).requires_grad_(True)
    got = ops.local_gcn(nd, labels.to(DEV), l1g, l2g, k, shift, norm, act, shortcut)
    (got * cot.to(DEV)).sum().backward()
    _close(got, want, 5e-5, "gcn out")

    def close_grad(a, b, what):
        # 2e-4 max-norm.  Fallback for ReLU flips: a unit whose pre-activation sits within fp32 rounding of zero takes the other
        # branch than the float64 reference; ONE such unit changes a whole row of its layer's weight gradient and, through
        # Adj^T (nearly uniform rows), every node gradient of the class and with them d_w1 -- a dense but small error
        # (measured: relative L2 1.6e-3 on all five gradients at once in a cosine / relu case, < 2e-5 otherwise): relative L2 <= 3e-3
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        scale = max(float(b.abs().max()), 1e-30)
        err = float((a - b).abs().max())
        if err <= 2e-4 * scale + 1e-7:
            return
        l2 = float((a - b).norm() / max(float(b.norm()), 1e-30))
        assert l2 <= 3e-3 and err <= 2e-2 * scale, "%s: max|d|=%.3e scale=%.3e relL2=%.3e" % (what, err, scale, l2)

    close_grad(nd.grad, nr.grad, "d_nodes")
    for name, a, b in (("w1", l1g.weight, l1d.weight), ("b1", l1g.bias, l1d.bias), ("w2", l2g.weight, l2d.weight), ("b2", l2g.bias, l2d.bias)):
        close_grad(a.grad, b.grad, "d_" + name)


@pytest.mark.parametrize("name", ["c8", "car", "dense"])
def test_fcos_loss_computation_matches_reference_golden(name, golden_dir):
    """f2: scan_fcos_assign_reg + the fused loss pass against vectors generated from the UNMODIFIED reference
    FCOSLossComputation (tests/tools/make_golden_fcos_loss.py): three losses and the gradients of all 15 head maps."""
    import os
    import fcos_loss_case
    from scan_b200 import fcos_hooks
    from scan_b200.config import scan_cfg
    from scan_b200.structures import BoxList
    gold = np.load(os.path.join(golden_dir, "fcos_loss.npz"))
    shapes, strides, boxes, labels, cls, reg, ctr, hw = fcos_loss_case.build(name)
    targets = []
    for b, l in zip(boxes, labels):
        t = BoxList(b, (hw[1], hw[0]), mode="xyxy")
        t.add_field("labels", l)
        targets.append(t)
    maps = [m.to(DEV).requires_grad_(True) for m in cls + reg + ctr]
    n = len(shapes)
    ev = fcos_hooks.make_fcos_loss_evaluator(scan_cfg("c2f"))
    losses = ev(None, maps[:n], maps[n:2 * n], maps[2 * n:], targets)
    (losses[0] * 1.0 + losses[1] * 0.7 + losses[2] * 1.3).backward()
    for i, v in enumerate(losses):
        w = float(gold["%s/loss%d" % (name, i)])
        assert abs(float(v) - w) <= 2e-5 * max(1.0, abs(w)), "loss %d: %r vs %r" % (i, float(v), w)
    for i, m in enumerate(maps):
        _close(m.grad, torch.from_numpy(gold["%s/grad%d" % (name, i)]), 2e-5, "grad %d" % i, atol=1e-9)
    # regression targets / labels bit-exact with the oracle's assignment
    geo = ops.Geometry(shapes, strides, len(boxes))
    pb, pl, pc, gmax = ops.pad_targets(targets, DEV)
    lab, rt = ops.fcos_assign_reg(geo, pb, pl, pc, gmax)
    wl, wr = orc.fcos_assign(shapes, strides, boxes, labels, return_reg=True)
    assert torch.equal(lab.cpu(), torch.cat(wl)) and torch.equal(rt.cpu(), torch.cat(wr))


def test_gather_rows_through_accumulates_into_the_downstream_gradient():
    g = torch.Generator().manual_seed(21)
    rows = torch.randn(3000, 256, generator=g)
    idx = torch.randint(0, 3000, (900,), generator=g).sort().values
    c1, c2 = torch.randn(900, 256, generator=g), torch.randn(3000, 256, generator=g)
    rr = rows.clone().requires_grad_(True)
    ((rr[idx] * c1).sum() + (rr * rr * c2).sum()).backward()
    rd = rows.to(DEV).requires_grad_(True)
    nodes, alias = ops.gather_rows_through(rd, idx.to(DEV).int())
    assert torch.equal(nodes.cpu(), rows[idx]) and torch.equal(alias, rd)
    ((nodes * c1.to(DEV)).sum() + (alias * alias * c2.to(DEV)).sum()).backward()
    _close(rd.grad, rr.grad, 1e-6, "d_rows (both consumers)")
    # only one of the two consumers takes part in the backward
    rd2 = rows.to(DEV).requires_grad_(True)
    n2, a2 = ops.gather_rows_through(rd2, idx.to(DEV).int())
    (n2 * c1.to(DEV)).sum().backward()
    _close(rd2.grad, torch.zeros_like(rows).index_add_(0, idx, c1), 1e-6, "d_rows (gather only)")
    rd3 = rows.to(DEV).requires_grad_(True)
    n3, a3 = ops.gather_rows_through(rd3, idx.to(DEV).int())
    (a3 * c2.to(DEV)).sum().backward()
    _close(rd3.grad, c2, 1e-6, "d_rows (alias only)")


@pytest.mark.parametrize("name", ["c8", "topk", "car_cap"])
def test_fcos_postprocessor_matches_reference_golden(name, golden_dir):
    """f4: candidate selection + top-k + decode + per-class NMS + detections cap (four launches) against vectors from the
    unmodified reference FCOSPostProcessor; detections compared as sets (label exact, score / box 1e-5), and in the
    reference's own order where that order is defined (no level above PRE_NMS_TOP_N)."""
    import os
    import postproc_case
    from scan_b200 import fcos_hooks
    gold = np.load(os.path.join(golden_dir, "postproc.npz"))
    d = postproc_case.build(name)
    pp = fcos_hooks.FCOSPostProcessor(pre_nms_thresh=d["thr"], pre_nms_top_n=d["top_n"], nms_thresh=d["nms_thr"],
                                      fpn_post_nms_top_n=d["post_n"], min_size=0, num_classes=d["num_fg"] + 1, mode="precision",
                                      fpn_strides=d["strides"])
    res = pp(None, [p.to(DEV) for p in d["probs"]], [r.to(DEV) for r in d["regs"]], [c.to(DEV) for c in d["ctrs"]], d["sizes"])
    assert len(res) == len(d["sizes"])
    for i, bl in enumerate(res):
        want = gold["%s/img%d" % (name, i)]
        b, s, l = bl.bbox.cpu().numpy(), bl.get_field("scores").cpu().numpy(), bl.get_field("labels").cpu().numpy()
        got = postproc_case.canonical(b, s, l)
        assert got.shape == want.shape, "image %d: %d detections, reference %d" % (i, got.shape[0], want.shape[0])
        assert np.array_equal(got[:, 0], want[:, 0])
        assert np.abs(got[:, 1:] - want[:, 1:]).max() <= 1e-5 * max(1.0, np.abs(want[:, 1:]).max())
        if name == "c8":
            assert np.array_equal(l, gold["%s/img%d_labels_in_order" % (name, i)])
            assert np.allclose(s, gold["%s/img%d_scores_in_order" % (name, i)], rtol=1e-6)
    # 'common' mode: logits in, the sigmoid is applied inside (inference.py:68)
    logits = [torch.logit(p.clamp(1e-6, 1 - 1e-6)).to(DEV) for p in d["probs"]]
    pp.mode = "common"
    res2 = pp(None, logits, [r.to(DEV) for r in d["regs"]], [c.to(DEV) for c in d["ctrs"]], d["sizes"])
    assert [len(a) for a in res2] == [len(a) for a in res]


def test_fcos_postprocessor_full_size_vs_oracle():
    """Full Cityscapes geometry, 8 images, PRE_NMS_TOP_N = 1000 reached at P3: against the CPU oracle restatement."""
    rs = np.random.RandomState(5)
    shapes, strides, n, c = FULL, STRIDES, 8, 8
    probs, regs, ctrs = [], [], []
    for (h, w), s in zip(shapes, strides):
        p = rs.uniform(0.0, 0.04, (n, c, h, w))
        hot = rs.rand(n, c, h, w) < 0.02
        p[hot] = rs.uniform(0.06, 0.99, int(hot.sum()))
        probs.append(torch.from_numpy(p.astype(np.float32)))
        regs.append(torch.from_numpy(np.exp(rs.standard_normal((n, 4, h, w)) * 0.5 + np.log(s * 2.0)).astype(np.float32)))
        ctrs.append(torch.from_numpy(rs.standard_normal((n, 1, h, w)).astype(np.float32)))
    sizes = [(800, 1344)] * n
    import postproc_case
    from scan_b200 import fcos_hooks
    want = orc.fcos_postprocess(shapes, strides, probs, regs, ctrs, sizes, 0.05, 1000, 0.6, 100)
    pp = fcos_hooks.FCOSPostProcessor(0.05, 1000, 0.6, 100, 0, c + 1, mode="light", fpn_strides=strides)
    res = pp(None, [p.to(DEV) for p in probs], [r.to(DEV) for r in regs], [t.to(DEV) for t in ctrs], sizes)
    for i, bl in enumerate(res):
        g = postproc_case.canonical(bl.bbox.cpu().numpy(), bl.get_field("scores").cpu().numpy(), bl.get_field("labels").cpu().numpy())
        w_ = postproc_case.canonical(want[i][0].numpy(), want[i][1].numpy(), want[i][2].numpy())
        assert g.shape == w_.shape and np.array_equal(g[:, 0], w_[:, 0])
        assert np.abs(g[:, 1:] - w_[:, 1:]).max() <= 1e-5 * max(1.0, np.abs(w_[:, 1:]).max())


def _conv_levels_ref(geo, x_rows, weight, bias):
    outs = []
    for l, (h, w) in enumerate(geo.shapes):
        x = x_rows[geo.row_off[l]:geo.row_off[l + 1]].view(geo.n_images, h, w, -1).permute(0, 3, 1, 2).double()
        y = torch.nn.functional.conv2d(x, weight.double(), None if bias is None else bias.double(), padding=1)
        outs.append(y.permute(0, 2, 3, 1).reshape(-1, weight.shape[0]))
    return torch.cat(outs)


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("shapes,n", [([(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)], 3), ([(100, 168), (50, 84)], 1), ([(9, 5)], 1)])
def test_conv3x3_rows_matches_fp64_conv(shapes, n, cta_group):
    """f1 forward / data-gradient kernel: 3xTF32 mode against an fp64 convolution at fp32 accuracy, single-pass TF32 mode at the
    TF32 bound (inputs rounded to 11 bits: 2^-11 relative per product, 2304 products)."""
    torch.manual_seed(3)
    geo = ops.Geometry(shapes, STRIDES[:len(shapes)], n)
    x = torch.randn(geo.R, 256, device=DEV)
    w = torch.randn(256, 256, 3, 3, device=DEV) * 0.02
    b = torch.randn(256, device=DEV)
    ref = _conv_levels_ref(geo, x, w, b)
    scale = float(ref.abs().max())
    for transpose in (False, True):
        wt = w if not transpose else w.flip(2, 3).transpose(0, 1)      # what the packed 'data gradient' weights compute
        r = ref if not transpose else _conv_levels_ref(geo, x, wt.contiguous(), b)
        hi, lo = ops.conv3x3_pack(w, transpose, True)
        y3 = ops.conv3x3_rows_raw(geo, x, hi, 256, bias=b, x_lo=ops.tf32_residual(x), packed_lo=lo, cta_group=cta_group)
        assert float((y3.double() - r).abs().max()) <= 5e-6 * scale, "3xTF32 %s" % transpose   # partial sums leave the truncating accumulator every 96 MMAs
        y1 = ops.conv3x3_rows_raw(geo, x, hi, 256, bias=b, cta_group=cta_group)
        assert float((y1.double() - r).abs().max()) <= 2e-3 * scale, "TF32 %s" % transpose
    # epilogue: addend + ReLU
    add = torch.randn(geo.R, 256, device=DEV)
    hi, lo = ops.conv3x3_pack(w, False, True)
    y = ops.conv3x3_rows_raw(geo, x, hi, 256, bias=b, addend=add, relu=True, x_lo=ops.tf32_residual(x), packed_lo=lo, cta_group=cta_group)
    assert float((y.double() - torch.relu(ref + add.double())).abs().max()) <= 5e-6 * scale


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("shapes,n", [([(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)], 3), ([(100, 168), (50, 84)], 2), ([(9, 5)], 1), ([(1, 1)], 1)])
def test_conv3x3_epilogue_group_norm_statistics(shapes, n, precise):
    """f1: the GroupNorm(32) statistics that come out of the tower convolution's epilogue (scan_conv3x3_rows_gn: per-warp partial
    sums over real pixels only, fp64 combination) against float64 statistics of the SAME output tensor + bias, and the
    normalised result against the kernels' own statistics pass; forward + backward through both operators."""
    torch.manual_seed(21)
    geo = ops.Geometry(shapes, STRIDES[:len(shapes)], n)
    xs = [torch.randn(n, 256, h, w, device=DEV).contiguous(memory_format=torch.channels_last) for h, w in shapes]
    w = (torch.randn(256, 256, 3, 3, device=DEV) * 0.02).requires_grad_(True)
    cb = (torch.randn(256, device=DEV) * 0.3).requires_grad_(True)
    gamma, beta = torch.randn(256, device=DEV).requires_grad_(True), (torch.randn(256, device=DEV) * 0.1).requires_grad_(True)
    saved = ops.CONV["precise"]
    ops.CONV["precise"] = precise
    try:
        res = {}
        for fused in (True, False):
            xi = [x.clone().requires_grad_(True) for x in xs]
            if fused:
                ys, stats = ops.conv3x3_levels(geo, w, xi, gn=(cb, 1e-5))
                assert stats.shape == (len(shapes) * n * 32 * 2,) and not stats.requires_grad
                for l, y in enumerate(ys):      # float64 statistics of the very tensor the kernel wrote
                    v = (y.detach().double() + cb.detach().double().view(1, -1, 1, 1)).reshape(n, 32, -1)
                    mean, var = v.mean(-1), v.var(-1, unbiased=False)
                    got = stats.view(len(shapes), n, 32, 2)[l].double()
                    scale = float(v.abs().max())
                    assert float((got[..., 0] - mean).abs().max()) <= 2e-6 * scale, "mean l%d" % l
                    rstd = 1.0 / torch.sqrt(var + 1e-5)
                    assert float(((got[..., 1] - rstd) / rstd).abs().max()) <= 2e-5, "rstd l%d" % l
            else:
                ys, stats = ops.conv3x3_levels(geo, w, xi), None
            hs = ops.gn_relu_levels(geo, gamma, beta, 1e-5, ys, conv_bias=cb, stats=stats)
            # d(loss)/dh vanishes at h = 0: a ReLU mask that flips on a last-bit difference of the statistics changes nothing
            loss = sum((h * h * torch.cos(h.detach() * 3.0)).sum() for h in hs)
            grads = torch.autograd.grad(loss, [w, cb, gamma, beta] + xi)
            res[fused] = ([h.detach() for h in hs], grads)
        for a, b in zip(res[True][0], res[False][0]):
            _close(a, b, 2e-5, "normalised output")
        # single-pass TF32: a last-bit difference of an input moves it across a truncation boundary of the tensor core's operand
        # read (2^-10 relative), so the two runs agree to the TF32 bound only; 3xTF32 runs agree to fp32 rounding
        for i, (a, b) in enumerate(zip(res[True][1], res[False][1])):
            _close(a, b, 1e-4 if precise else 2e-3, "gradient %d" % i)
    finally:
        ops.CONV["precise"] = saved


@pytest.mark.parametrize("shapes,n", [([(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)], 3), ([(100, 168), (50, 84)], 2), ([(9, 5)], 1)])
def test_conv3x3_wgrad_matches_fp64(shapes, n):
    """f1 weight-gradient kernel (MN-major operands, CTA pairs, deterministic split over pixel segments) against the fp64
    gradient of torch's convolution: 3xTF32 at fp32 accuracy, single-pass TF32 at the TF32 bound; bitwise repeatable."""
    torch.manual_seed(4)
    geo = ops.Geometry(shapes, STRIDES[:len(shapes)], n)
    x = torch.randn(geo.R, 256, device=DEV)
    dy = torch.randn(geo.R, 256, device=DEV)
    w = torch.zeros(256, 256, 3, 3, device=DEV, dtype=torch.float64, requires_grad=True)
    ys = []
    for l, (h, wd) in enumerate(geo.shapes):
        xl = x[geo.row_off[l]:geo.row_off[l + 1]].view(n, h, wd, 256).permute(0, 3, 1, 2).double()
        ys.append(torch.nn.functional.conv2d(xl, w, None, padding=1).permute(0, 2, 3, 1).reshape(-1, 256))
    (ref,) = torch.autograd.grad(torch.cat(ys), [w], dy.double())
    scale = float(ref.abs().max())
    g3 = ops.conv3x3_wgrad_raw(geo, x, dy, x_lo=ops.tf32_residual(x), dy_lo=ops.tf32_residual(dy))
    assert float((g3.double() - ref).abs().max()) <= 5e-6 * scale
    g1 = ops.conv3x3_wgrad_raw(geo, x, dy)
    assert float((g1.double() - ref).abs().max()) <= 2e-3 * scale
    assert torch.equal(g1, ops.conv3x3_wgrad_raw(geo, x, dy))
    # a channels-last gradient tensor (what torch hands out for channels-last weights)
    out = torch.empty(256, 256, 3, 3, device=DEV).contiguous(memory_format=torch.channels_last)
    ops.conv3x3_wgrad_raw(geo, x, dy, out=out)
    assert torch.equal(out, g1)


@pytest.mark.parametrize("name", ["c9_source", "c9_target", "k2_target_only"])
def test_cka_discriminator_matches_reference_golden(name, golden_dir):
    """f3: FCOSDiscriminator_con (dis_tower, all-classes conditional maps as two tcgen05 convolutions, weighted BCE, gradient
    reversal into the features AND the activation maps) against vectors of the unmodified reference module: loss, d(feature),
    d(act_maps), every parameter gradient.  3xTF32 convolutions (parity mode), bound 1e-3 of each tensor's max (measured ~1e-5)."""
    import os
    import cka_case
    from scan_b200.discriminator import FCOSDiscriminator_con
    gold = np.load(os.path.join(golden_dir, "cka.npz"))
    d = cka_case.build(name)
    m = FCOSDiscriminator_con(num_convs=d["num_convs"], num_classes=d["k"], grad_reverse_lambda=d["lam"], grl_applied_domain=d["grl_dom"])
    m.load_state_dict(cka_case.state_dict_for(m, seed=5))
    m.to(DEV)
    saved = ops.CONV["precise"]
    ops.CONV["precise"] = True
    try:
        feat = d["feat"].to(DEV).requires_grad_(True)
        act = d["act"].to(DEV).requires_grad_(True)
        loss = m(feat, d["target"], act_maps=act, domain=d["domain"])
        loss.backward()
    finally:
        ops.CONV["precise"] = saved
    want = float(gold[name + "/loss"])
    assert abs(float(loss.detach()) - want) <= 2e-5 * max(1.0, abs(want)), (float(loss.detach()), want)
    got = {"d_feature": feat.grad, "d_act": act.grad if act.grad is not None else torch.zeros_like(act)}
    for k, p in m.named_parameters():
        got["grad/" + k] = p.grad
    worst = 0.0
    for k, v in got.items():
        worst = max(worst, cka_case.check(v.detach().cpu().numpy(), gold[name + "/" + k + "#s"], gold[name + "/" + k + "#n"], 1e-3))
    print("cka %s worst relative-to-max error %.2e" % (name, worst))
    # the benchmark arithmetic (single-pass TF32 convolutions): same loss within the TF32 bound
    loss_fast = m(d["feat"].to(DEV), d["target"], act_maps=d["act"].to(DEV), domain=d["domain"])
    assert abs(float(loss_fast) - want) <= 5e-3 * max(1.0, abs(want))


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("k", [9, 2])
def test_head_out_levels_matches_fp64(fused, k):
    """head_out on the tower kernels (two-input convolution with bias + ReLU epilogue; N = 32 data gradient for the maps): output
    and every gradient against torch's fp64 conv2d(cat([features, maps], 1)); 3xTF32 mode, 1e-4 of each tensor's max."""
    torch.manual_seed(6)
    shapes, n = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)], 2
    geo = ops.Geometry(shapes, STRIDES, n)
    feats = [torch.randn(n, 256, h, w, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for h, w in shapes]
    acts = [torch.softmax(torch.randn(n, k, h, w, device=DEV), 1).requires_grad_(True) for h, w in shapes]
    weight = (torch.randn(256, 256 + k, 3, 3, device=DEV) * 0.03).requires_grad_(True)
    bias = (torch.randn(256, device=DEV) * 0.1).requires_grad_(True)
    cots = [torch.randn(n, 256, h, w, device=DEV) for h, w in shapes]
    saved = ops.CONV["precise"]
    ops.CONV["precise"] = True
    try:
        if fused:
            ys = ops.head_out_levels(geo, weight, bias, acts, features=feats)
        else:
            us = ops.conv3x3_levels(geo, weight[:, :256], feats)
            ys = ops.head_out_levels(geo, weight, bias, acts, us=us)
        torch.autograd.backward(ys, cots)
    finally:
        ops.CONV["precise"] = saved
    got = [y.detach() for y in ys] + [f.grad for f in feats] + [a.grad for a in acts] + [weight.grad, bias.grad]
    f64 = [f.detach().double().requires_grad_(True) for f in feats]
    a64 = [a.detach().double().requires_grad_(True) for a in acts]
    w64, b64 = weight.detach().double().requires_grad_(True), bias.detach().double().requires_grad_(True)
    y64 = [torch.relu(torch.nn.functional.conv2d(torch.cat([f, a], 1), w64, b64, padding=1)) for f, a in zip(f64, a64)]
    torch.autograd.backward(y64, [c.double() for c in cots])
    want = [y.detach() for y in y64] + [f.grad for f in f64] + [a.grad for a in a64] + [w64.grad, b64.grad]
    for i, (g, w_) in enumerate(zip(got, want)):
        scale = float(w_.abs().max())
        assert float((g.double() - w_).abs().max()) <= 1e-4 * scale, (i, float((g.double() - w_).abs().max()), scale)


def test_alias_chain_condconv_head_out_gather_matches_plain_autograd():
    """The alias chain of the training branches (conditional conv hands the rows on, head_out reads them as one matrix and hands
    them on again, the node gather comes last): every gradient must equal the plain formulation in which autograd sums the three
    d_rows contributions itself (same kernels otherwise, so the comparison is tight: 1e-5 of each tensor's max)."""
    torch.manual_seed(8)
    shapes, n, k = [(25, 42), (13, 21), (7, 11), (4, 6), (2, 3)], 2, 9
    geo = ops.Geometry(shapes, STRIDES, n)
    base = torch.randn(geo.R, 256, device=DEV)
    w_cc = torch.randn(k, 256, device=DEV) * 0.05
    w_ho = torch.randn(256, 256 + k, 3, 3, device=DEV) * 0.03
    b_ho = torch.randn(256, device=DEV) * 0.1
    idx = torch.randperm(geo.R, device=DEV)[:300].sort().values.to(torch.int32)
    cot_y = [torch.randn(n, 256, h, w, device=DEV) for h, w in shapes]
    cot_a = [torch.randn(n, k, h, w, device=DEV) * 0.1 for h, w in shapes]
    cot_n = torch.randn(300, 256, device=DEV)
    saved = ops.CONV["precise"]
    ops.CONV["precise"] = True
    results = []
    try:
        for chain in (True, False):
            rows = base.clone().requires_grad_(True)
            wc, wh, bh = (t.clone().requires_grad_(True) for t in (w_cc, w_ho, b_ho))
            if chain:
                acts, _, _, r1 = ops.condconv(geo, rows, wc, None, k, 0, through=True)
                ys, r2 = ops.head_out_levels(geo, wh, bh, acts, rows=r1, through=True)
                nodes = ops.gather_rows(r2, idx)
            else:
                acts, _, _ = ops.condconv(geo, rows, wc, None, k, 0)
                ys = ops.head_out_levels(geo, wh, bh, acts, features=ops.level_views(geo, rows))
                nodes = ops.gather_rows(rows, idx)
            torch.autograd.backward(list(ys) + list(acts) + [nodes], cot_y + cot_a + [cot_n])
            results.append([y.detach() for y in ys] + [rows.grad, wc.grad, wh.grad, bh.grad])
    finally:
        ops.CONV["precise"] = saved
    for i, (a, b) in enumerate(zip(*results)):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-5 * scale, (i, float((a - b).abs().max()), scale)
