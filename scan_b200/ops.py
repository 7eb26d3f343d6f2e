"""Host-side operators: torch tensors in, C-ABI calls on the current CUDA stream, torch tensors out.

torch is plumbing here (device memory, streams, autograd bookkeeping); the arithmetic happens in
libscan_b200.so.  Every op raises if its input is not a CUDA tensor: there is no CPU fallback.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ScanLevels, ScanSampleMeta

C = 256  # channel width of the middle head (PROTO_CHANNEL / FPN channels)

# Optional per-entry-point device timing (bench.py): CUDA events on the launching stream around every ABI call.
TIMING = {"on": False}
TIMERS = []


def call(name, *args):
    if TIMING["on"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call(name, *args)
        e1.record()
        TIMERS.append((name, e0, e1))
    else:
        _lib.call(name, *args)


def timers_summary():
    """{entry point: {"ms": total device ms, "calls": n}} of the calls recorded since TIMERS.clear()."""
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in TIMERS:
        d = out.setdefault(name.replace("scan_", "", 1), {"ms": 0.0, "calls": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["calls"] += 1
    return out


def _stream():
    # raw handle of torch's current stream (torch.cuda.current_stream() builds a Stream object: ~17 us a call, ~100 calls a step)
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("scan_b200 ops need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def _ptr_array(tensors, n=8):
    arr = (ctypes.c_void_p * n)()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class Geometry(object):
    """FPN level shapes of one call: the `rows` layout of include/scan_b200.h."""

    def __init__(self, shapes, strides, n_images):
        if len(shapes) > _lib.SCAN_MAX_LEVELS:
            raise RuntimeError("at most %d FPN levels" % _lib.SCAN_MAX_LEVELS)
        self.shapes = [(int(h), int(w)) for h, w in shapes]
        self.strides = [int(s) for s in strides][:len(shapes)]
        self.n_images = int(n_images)
        lv = ScanLevels()
        lv.n_levels = len(shapes)
        lv.n_images = self.n_images
        off = [0]
        for l, (h, w) in enumerate(self.shapes):
            lv.h[l], lv.w[l], lv.stride[l] = h, w, self.strides[l]
            off.append(off[-1] + self.n_images * h * w)
        self.levels = lv
        self.row_off = off
        self.R = off[-1]

    def ref(self):
        return ctypes.byref(self.levels)

    def split_rows(self, t):
        """[R, ...] tensor -> list of per-level views."""
        return [t[self.row_off[l]:self.row_off[l + 1]] for l in range(len(self.shapes))]

    @classmethod
    def of(cls, features, strides):
        return cls([tuple(f.shape[-2:]) for f in features], strides, features[0].shape[0])


# ----------------------------------------------------------------------------------------------------
# layout
# ----------------------------------------------------------------------------------------------------
class _PackRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geo, *feats):
        feats = [f.contiguous() for f in feats]
        rows = torch.empty((geo.R, C), device=feats[0].device, dtype=torch.float32)
        call("scan_pack_rows", geo.ref(), _ptr_array(feats), C, _ptr(rows), _stream())
        ctx.geo = geo
        ctx.shapes = [f.shape for f in feats]
        return rows

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_rows):
        d_rows = d_rows.contiguous()
        grads = [torch.empty(s, device=d_rows.device, dtype=torch.float32) for s in ctx.shapes]
        call("scan_unpack_rows", ctx.geo.ref(), _ptr(d_rows), C, _ptr_array(grads), 0, _stream())
        return (None,) + tuple(grads)


def pack_rows(geo, feats):
    for f in feats:
        if f.dtype != torch.float32 or f.shape[1] != C:
            raise RuntimeError("features must be fp32 with %d channels" % C)
    return _PackRows.apply(geo, *feats)


# ----------------------------------------------------------------------------------------------------
# rows <-> per-level channels-last tensors (zero-copy), towers on the rows layout
# ----------------------------------------------------------------------------------------------------
def level_views(geo, rows):
    """[R,256] rows -> list of [N,256,H_l,W_l] tensors that are channels_last VIEWS of `rows` (no copy)."""
    return [rows[geo.row_off[l]:geo.row_off[l + 1]].view(geo.n_images, h, w, C).permute(0, 3, 1, 2)
            for l, (h, w) in enumerate(geo.shapes)]


def nhwc_dense(t):
    """[N,C,H,W] tensor whose memory is NHWC-dense (torch channels_last); copies only if it is not already."""
    if t.permute(0, 2, 3, 1).is_contiguous():
        return t
    return t.contiguous(memory_format=torch.channels_last)


def _level_rows(t):
    """[N,C,H,W] NHWC-dense tensor -> its [N*H*W, C] rows view."""
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])


class _PackLevels(torch.autograd.Function):
    """NCHW features -> the rows matrix, handed out as per-level channels_last views (what cuDNN's NHWC kernels want)."""

    @staticmethod
    def forward(ctx, geo, *feats):
        feats = [f.contiguous() for f in feats]
        rows = torch.empty((geo.R, C), device=feats[0].device, dtype=torch.float32)
        call("scan_pack_rows", geo.ref(), _ptr_array(feats), C, _ptr(rows), _stream())
        ctx.geo = geo
        ctx.shapes = [f.shape for f in feats]
        return tuple(level_views(geo, rows))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *d_levels):
        geo = ctx.geo
        if any(g is None for g in d_levels):     # a level without gradient: per-level launches for the others
            grads = []
            for l, (g, shape) in enumerate(zip(d_levels, ctx.shapes)):
                if g is None:
                    grads.append(None)
                    continue
                g_rows = _level_rows(nhwc_dense(g))
                out = torch.empty(shape, device=g.device, dtype=torch.float32)
                one = Geometry([geo.shapes[l]], [geo.strides[l]], geo.n_images)
                call("scan_unpack_rows", one.ref(), _ptr(g_rows), C, _ptr_array([out]), 0, _stream())
                grads.append(out)
            return (None,) + tuple(grads)
        # all five levels in ONE launch, straight from the per-level gradient tensors
        g_rows = [_level_rows(nhwc_dense(g)) for g in d_levels]
        grads = [torch.empty(shape, device=g_rows[0].device, dtype=torch.float32) for shape in ctx.shapes]
        call("scan_unpack_levels", geo.ref(), _ptr_array(g_rows), C, _ptr_array(grads), _stream())
        return (None,) + tuple(grads)


def pack_levels(geo, feats):
    for f in feats:
        if f.dtype != torch.float32 or f.shape[1] != C:
            raise RuntimeError("features must be fp32 with %d channels" % C)
    return list(_PackLevels.apply(geo, *feats))


class _JoinRows(torch.autograd.Function):
    """Per-level tensors -> [R,256] rows.  Zero-copy when the levels are adjacent channels_last views of one buffer (what
    gn_relu_levels / pack_levels hand out); otherwise one concatenating copy."""

    @staticmethod
    def forward(ctx, geo, *levels):
        ctx.geo = geo
        base = levels[0]
        adjacent = True
        for l, t in enumerate(levels):
            if (not t.permute(0, 2, 3, 1).is_contiguous() or t.untyped_storage().data_ptr() != base.untyped_storage().data_ptr()
                    or t.data_ptr() != base.data_ptr() + geo.row_off[l] * C * 4):
                adjacent = False
        if adjacent:
            return torch.as_strided(base.detach(), (geo.R, C), (C, 1), base.storage_offset())
        return torch.cat([_level_rows(nhwc_dense(t)) for t in levels])

    @staticmethod
    def backward(ctx, d_rows):
        return (None,) + tuple(level_views(ctx.geo, d_rows.contiguous()))


def join_rows(geo, levels):
    return _JoinRows.apply(geo, *levels)


class _GnReluLevels(torch.autograd.Function):
    """relu(group_norm(x + conv_bias, 32)) of the five per-level convolution outputs in one launch sequence; the result is
    ONE rows buffer handed out as per-level channels_last views.  conv_bias (optional) is the bias of the convolution that
    produced x: folded in here so that the convolution runs bias-free and its bias gradient is a by-product."""

    @staticmethod
    def forward(ctx, geo, gamma, beta, conv_bias, eps, stats, *xs):
        xs = [nhwc_dense(x) for x in xs]
        dev = xs[0].device
        gamma, beta = gamma.contiguous(), beta.contiguous()
        conv_bias = None if conv_bias is None else conv_bias.contiguous()
        y_rows = torch.empty((geo.R, C), device=dev, dtype=torch.float32)
        if stats is not None:     # statistics from the convolution's epilogue (conv3x3_levels(..., gn=...)): apply pass only
            call("scan_gn_relu_apply", geo.ref(), _ptr_array(xs), _ptr(conv_bias), _ptr(gamma), _ptr(beta), _ptr(stats), _ptr(y_rows),
                 _stream())
        else:
            stats = torch.empty((len(xs) * geo.n_images * 32 * 2,), device=dev, dtype=torch.float32)
            ws = torch.empty((_lib.lib().scan_gn_workspace_bytes(geo.ref()),), device=dev, dtype=torch.uint8)
            call("scan_gn_relu_fwd", geo.ref(), _ptr_array(xs), _ptr(conv_bias), _ptr(gamma), _ptr(beta), float(eps), _ptr(y_rows),
                 _ptr(stats), _ptr(ws), ws.numel(), _stream())
        ctx.geo = geo
        ctx.has_cbias = conv_bias is not None
        ctx.save_for_backward(gamma, stats, beta, gamma if conv_bias is None else conv_bias, *xs)
        return tuple(level_views(geo, y_rows))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *d_levels):
        geo = ctx.geo
        gamma, stats, beta, cbias = ctx.saved_tensors[:4]
        xs = ctx.saved_tensors[4:]
        dev = gamma.device
        if not ctx.has_cbias:
            cbias = None
        dys = [nhwc_dense(g) if g is not None else torch.zeros_like(x) for g, x in zip(d_levels, xs)]
        dx_rows = torch.empty((geo.R, C), device=dev, dtype=torch.float32)     # y is not read: the kernels recompute the ReLU mask
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        dcb = torch.empty_like(gamma) if cbias is not None else None
        ws = torch.empty((_lib.lib().scan_gn_workspace_bytes(geo.ref()),), device=dev, dtype=torch.uint8)
        call("scan_gn_relu_bwd", geo.ref(), _ptr_array(xs), _ptr_array(dys), _ptr(cbias), _ptr(gamma), _ptr(beta), _ptr(stats),
             _ptr(dx_rows), _ptr(dgamma), _ptr(dbeta), _ptr(dcb), _ptr(ws), ws.numel(), _stream())
        return (None, dgamma, dbeta, dcb, None, None) + tuple(level_views(geo, dx_rows))


def gn_relu_levels(geo, gamma, beta, eps, xs, conv_bias=None, stats=None):
    """stats: the [L*N*32, 2] (mean, rstd) array conv3x3_levels(..., gn=(conv_bias, eps)) returned for these xs, or None (the
    kernels then take a statistics pass of their own)."""
    if gamma.numel() != C:
        raise RuntimeError("the GroupNorm kernel is built for %d channels in 32 groups" % C)
    for x in xs:
        if not x.is_cuda or x.dtype != torch.float32 or x.shape[1] != C:
            raise RuntimeError("gn_relu_levels expects CUDA fp32 [N,%d,H,W] tensors (no CPU fallback)" % C)
    return list(_GnReluLevels.apply(geo, gamma, beta, conv_bias, eps, stats, *xs))


# ----------------------------------------------------------------------------------------------------
# f1: tower convolutions as a tcgen05 implicit GEMM on the rows layout (csrc/tower.cu)
# ----------------------------------------------------------------------------------------------------
CONV = {"cta_group": 2, "precise": False,     # precise = 3xTF32 (fp32-accurate; the parity runs), else single-pass TF32
        # GroupNorm statistics as a by-product of the tower convolution's epilogue (SCAN_B200_GN_STATS=0: separate statistics pass)
        "gn_stats": os.environ.get("SCAN_B200_GN_STATS", "1") != "0"}


_WORKSPACES = {}


def _workspace(tag, nbytes, device):
    """A per-(tag, device, stream) scratch buffer that is reused across calls (stream order serialises its users): the ~100 MB
    split-K partials of the weight gradient would otherwise go through the caching allocator three times per step."""
    key = (tag, torch.device(device).index, torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty((nbytes,), device=device, dtype=torch.uint8)
        _WORKSPACES[key] = ws
    return ws


def conv3x3_pack(weight, transpose, precise):
    """[Cout,Cin,3,3] weight (any strides) -> (packed_hi, packed_lo | None) for scan_conv3x3_rows."""
    cout, cin = weight.shape[0], weight.shape[1]
    rows, cols = (cin, cout) if transpose else (cout, cin)
    n = _lib.lib().scan_conv3x3_packed_floats(rows, cols)
    hi = torch.empty((n,), device=weight.device, dtype=torch.float32)
    lo = torch.empty_like(hi) if precise else None
    s = weight.stride()
    call("scan_conv3x3_pack_weights", _ptr(weight), s[0], s[1], s[2], s[3], cout, cin, int(bool(transpose)), _ptr(hi), _ptr(lo),
         _stream())
    return hi, lo


def tf32_residual(x):
    lo = torch.empty_like(x)
    call("scan_tf32_residual", _ptr(x), x.numel(), _ptr(lo), _stream())
    return lo


def conv3x3_rows_raw(geo, x_rows, packed, n_out, bias=None, addend=None, relu=False, x_lo=None, packed_lo=None, out=None, cta_group=None):
    """out[R, n_out] = act(conv3x3(x_rows[R, cin]) + bias + addend) with pre-packed weights (no autograd)."""
    if not x_rows.is_cuda or x_rows.dtype != torch.float32 or not x_rows.is_contiguous() or x_rows.shape[0] != geo.R:
        raise RuntimeError("conv3x3_rows expects a contiguous CUDA fp32 [R, Cin] rows matrix (no CPU fallback)")
    if out is None:
        out = torch.empty((geo.R, n_out), device=x_rows.device, dtype=torch.float32)
    call("scan_conv3x3_rows", geo.ref(), _ptr(x_rows), _ptr(x_lo), x_rows.shape[1], _ptr(packed), _ptr(packed_lo), n_out, _ptr(bias),
         _ptr(addend), int(bool(relu)), _ptr(out), out.shape[1], CONV["cta_group"] if cta_group is None else cta_group, _stream())
    return out


def conv3x3_wgrad_raw(geo, x_rows, dy_rows, x_lo=None, dy_lo=None, out=None):
    """d_w [Cout,Cin,3,3] of the tower convolution from the rows matrices (no autograd)."""
    cin, cout = x_rows.shape[1], dy_rows.shape[1]
    if out is None:
        out = torch.empty((cout, cin, 3, 3), device=x_rows.device, dtype=torch.float32)
    nbytes = _lib.lib().scan_conv3x3_wgrad_workspace_bytes(geo.ref(), cin, cout, int(x_lo is not None))
    if nbytes < 0:
        raise RuntimeError("conv3x3_wgrad: channel counts must be multiples of 256")
    ws = _workspace("conv_wgrad", nbytes, x_rows.device)
    s = out.stride()
    call("scan_conv3x3_wgrad", geo.ref(), _ptr(x_rows), _ptr(x_lo), cin, _ptr(dy_rows), _ptr(dy_lo), cout, _ptr(out), s[0], s[1], s[2],
         s[3], _ptr(ws), nbytes, _stream())
    return out


def _rows_of_levels(geo, levels):
    """Per-level [N,C,H,W] tensors -> [R,C] rows; zero-copy when they are adjacent channels_last views of one buffer."""
    base = levels[0]
    c = base.shape[1]
    for l, t in enumerate(levels):
        if (not t.permute(0, 2, 3, 1).is_contiguous() or t.untyped_storage().data_ptr() != base.untyped_storage().data_ptr()
                or t.data_ptr() != base.data_ptr() + geo.row_off[l] * c * 4):
            return torch.cat([_level_rows(nhwc_dense(t)) for t in levels])
    return torch.as_strided(base.detach(), (geo.R, c), (c, 1), base.storage_offset())


def _level_views_c(geo, rows):
    c = rows.shape[1]
    return [rows[geo.row_off[l]:geo.row_off[l + 1]].view(geo.n_images, h, w, c).permute(0, 3, 1, 2) for l, (h, w) in enumerate(geo.shapes)]


class _Conv3x3Levels(torch.autograd.Function):
    """The bias-free 3x3 convolution (padding 1) of a tower layer over all levels: ONE scan_conv3x3_rows launch forward, one for
    the data gradient (the same kernel on dY with the rotated, transposed weights) and scan_conv3x3_wgrad for the weights.
    Inputs / outputs are per-level channels_last views of rows buffers."""

    @staticmethod
    def forward(ctx, geo, weight, gn, *levels):
        """gn = None | (conv_bias | None, eps): also return the GroupNorm(32) statistics of (y + conv_bias) as a last,
        non-differentiable output (the dependence of the statistics on y is part of _GnReluLevels' backward formula)."""
        precise = CONV["precise"]
        x_rows = _rows_of_levels(geo, levels)
        hi, lo = conv3x3_pack(weight, False, precise)
        x_lo = tf32_residual(x_rows) if precise else None
        ctx.geo, ctx.precise, ctx.with_stats = geo, precise, gn is not None
        ctx.save_for_backward(x_rows, weight)
        ctx.x_lo = x_lo
        if gn is None:
            y_rows = conv3x3_rows_raw(geo, x_rows, hi, weight.shape[0], x_lo=x_lo, packed_lo=lo)
            return tuple(_level_views_c(geo, y_rows))
        gn_bias, eps = gn
        gn_bias = None if gn_bias is None else gn_bias.detach().contiguous()
        y_rows = torch.empty((geo.R, C), device=x_rows.device, dtype=torch.float32)
        stats = torch.empty((len(levels) * geo.n_images * 32 * 2,), device=x_rows.device, dtype=torch.float32)
        nbytes = _lib.lib().scan_conv3x3_gn_workspace_bytes(geo.ref())
        ws = _workspace("conv_gn", nbytes, x_rows.device)
        call("scan_conv3x3_rows_gn", geo.ref(), _ptr(x_rows), _ptr(x_lo), x_rows.shape[1], _ptr(hi), _ptr(lo), _ptr(gn_bias), float(eps),
             _ptr(y_rows), _ptr(stats), CONV["cta_group"], _ptr(ws), nbytes, _stream())
        ctx.mark_non_differentiable(stats)
        return tuple(_level_views_c(geo, y_rows)) + (stats,)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *d_levels):
        geo, precise = ctx.geo, ctx.precise
        if ctx.with_stats:
            d_levels = d_levels[:-1]
        x_rows, weight = ctx.saved_tensors
        d_levels = [g if g is not None else torch.zeros((geo.n_images, weight.shape[0], h, w), device=x_rows.device).contiguous(
            memory_format=torch.channels_last) for g, (h, w) in zip(d_levels, geo.shapes)]
        dy_rows = _rows_of_levels(geo, d_levels)
        dy_lo = tf32_residual(dy_rows) if precise else None
        d_w = None
        if ctx.needs_input_grad[1]:
            x_lo = ctx.x_lo if precise else None
            d_w = torch.empty_like(weight)
            conv3x3_wgrad_raw(geo, x_rows, dy_rows, x_lo=x_lo, dy_lo=dy_lo, out=d_w)
        d_x = (None,) * len(d_levels)
        if any(ctx.needs_input_grad[2:]):
            hi, lo = conv3x3_pack(weight, True, precise)
            dx_rows = conv3x3_rows_raw(geo, dy_rows, hi, weight.shape[1], x_lo=dy_lo, packed_lo=lo)
            d_x = tuple(_level_views_c(geo, dx_rows))
        return (None, d_w, None) + d_x


def conv3x3_levels(geo, weight, levels, gn=None):
    """Tower convolution (no bias) of per-level [N,Cin,H,W] CUDA tensors -> per-level channels_last views of one rows buffer.
    gn = (conv_bias | None, eps) and 256 output channels: returns (levels, stats) with the GroupNorm(32) statistics of
    (y + conv_bias) for gn_relu_levels(..., stats=stats)."""
    if weight.shape[2:] != (3, 3) or weight.shape[0] % 256 or weight.shape[1] % 256:
        raise RuntimeError("conv3x3_levels is built for 3x3 kernels with channel counts in multiples of 256")
    for x in levels:
        if not x.is_cuda or x.dtype != torch.float32 or x.shape[1] != weight.shape[1]:
            raise RuntimeError("conv3x3_levels expects CUDA fp32 [N,%d,H,W] tensors (no CPU fallback)" % weight.shape[1])
    if gn is not None:
        if weight.shape[0] != C:
            raise RuntimeError("GroupNorm statistics from the convolution epilogue need %d output channels" % C)
        out = _Conv3x3Levels.apply(geo, weight, gn, *levels)
        return list(out[:-1]), out[-1]
    return list(_Conv3x3Levels.apply(geo, weight, None, *levels))


class _AddReluLevels(torch.autograd.Function):
    """relu(u + v + bias) per level (head_out without the concat), all levels per launch, output = views of one rows buffer."""

    @staticmethod
    def forward(ctx, geo, bias, n_levels, *uv):
        us = [nhwc_dense(t) for t in uv[:n_levels]]
        vs = [nhwc_dense(t) for t in uv[n_levels:]]
        dev = us[0].device
        bias = None if bias is None else bias.contiguous()
        y_rows = torch.empty((geo.R, C), device=dev, dtype=torch.float32)
        call("scan_add_relu_fwd", geo.ref(), _ptr_array(us), _ptr_array(vs) if vs else None, _ptr(bias), _ptr(y_rows), _stream())
        ctx.geo = geo
        ctx.has_bias = bias is not None
        ctx.n_levels, ctx.has_v = n_levels, bool(vs)
        ctx.save_for_backward(y_rows)
        return tuple(level_views(geo, y_rows))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *d_levels):
        geo = ctx.geo
        (y_rows,) = ctx.saved_tensors
        dev = y_rows.device
        views = level_views(geo, y_rows)
        dys = [nhwc_dense(g) if g is not None else torch.zeros_like(v) for g, v in zip(d_levels, views)]
        d_rows = torch.empty_like(y_rows)
        d_bias = torch.empty((C,), device=dev, dtype=torch.float32) if ctx.has_bias else None
        ws = torch.empty((_lib.lib().scan_gn_workspace_bytes(geo.ref()),), device=dev, dtype=torch.uint8) if ctx.has_bias else None
        call("scan_add_relu_bwd", geo.ref(), _ptr_array(dys), _ptr(y_rows), _ptr(d_rows), _ptr(d_bias), _ptr(ws),
             0 if ws is None else ws.numel(), _stream())
        d = tuple(level_views(geo, d_rows))
        return (None, d_bias, None) + d + (d if ctx.has_v else ())


def add_relu_levels(geo, bias, us, vs=None):
    for x in list(us) + list(vs or []):
        if not x.is_cuda or x.dtype != torch.float32 or x.shape[1] != C:
            raise RuntimeError("add_relu_levels expects CUDA fp32 [N,%d,H,W] tensors (no CPU fallback)" % C)
    return list(_AddReluLevels.apply(geo, bias, len(us), *(list(us) + list(vs or []))))


# ----------------------------------------------------------------------------------------------------
# K1: assignment, sampling, gather
# ----------------------------------------------------------------------------------------------------
def pad_targets(targets, device):
    """list[BoxList] -> (boxes [N,G,4] fp32, labels [N,G] int64, count [N] int32) on `device`."""
    n = len(targets)
    counts = []
    for t in targets:
        if t.mode != "xyxy":
            raise AssertionError("targets must be in xyxy mode (loss.py:308)")
        counts.append(int(t.bbox.shape[0]))
    if min(counts) == 0:
        raise RuntimeError("an image without ground-truth boxes is unsupported (the reference's empty min, loss.py:333)")
    g = max(counts)
    if all(t.bbox.is_cuda for t in targets):
        # targets already live on the GPU (the reference's trainer moves them there): pad with device copies, no D2H sync
        boxes = torch.zeros((n, g, 4), device=device, dtype=torch.float32)
        labels = torch.zeros((n, g), device=device, dtype=torch.int64)
        for i, t in enumerate(targets):
            boxes[i, :counts[i]] = t.bbox.detach().to(device, torch.float32)
            labels[i, :counts[i]] = t.get_field("labels").detach().to(device, torch.int64)
        cnt = _upload_small(torch.tensor(counts, dtype=torch.int32), device)
        return boxes, labels, cnt, g
    # host targets: ONE pinned staging buffer (boxes | labels | counts; the kernel reads boxes as float4: 16-byte aligned first),
    # read by a kernel -- not the copy engine (see _upload_small)
    nb_box, nb_lab = n * g * 16, n * g * 8
    stage = torch.zeros((nb_box + nb_lab + n * 4,), dtype=torch.uint8, pin_memory=True)
    boxes_h = stage[:nb_box].view(torch.float32).view(n, g, 4)
    labels_h = stage[nb_box:nb_box + nb_lab].view(torch.int64).view(n, g)
    cnt_h = stage[nb_box + nb_lab:].view(torch.int32)
    for i, t in enumerate(targets):
        boxes_h[i, :counts[i]] = t.bbox.detach().to("cpu", torch.float32)
        labels_h[i, :counts[i]] = t.get_field("labels").detach().to("cpu", torch.int64)
    cnt_h.copy_(torch.tensor(counts, dtype=torch.int32))
    dev_buf = _upload_small(stage, device)
    return (dev_buf[:nb_box].view(torch.float32).view(n, g, 4), dev_buf[nb_box:nb_box + nb_lab].view(torch.int64).view(n, g),
            dev_buf[nb_box + nb_lab:].view(torch.int32), g)


_UPLOADS = []      # (event, pinned buffer): keeps a staging buffer alive until the stream has consumed it


def _upload_small(host, device):
    """Small host tensor -> device through scan_upload_small: the kernel reads the PINNED host buffer directly (unified
    addressing), so the transfer does not queue on the copy engine behind a multi-hundred-MB input prefetch of the training loop
    (measured: with the batch of the next step in flight, the source pass' first host read waited 6 ms for a 2 KB copy)."""
    src = host.contiguous()
    if not src.is_pinned():
        src = src.pin_memory()
    nbytes = src.numel() * src.element_size()
    pad = (-nbytes) % 4
    if pad:
        raise RuntimeError("_upload_small expects a multiple of 4 bytes")
    out = torch.empty(src.shape, device=device, dtype=src.dtype)
    call("scan_upload_small", ctypes.c_void_p(src.data_ptr()), _ptr(out), nbytes, _stream())
    ev = torch.cuda.Event()
    ev.record()
    _UPLOADS.append((ev, src))
    while len(_UPLOADS) > 4 and _UPLOADS[0][0].query():
        _UPLOADS.pop(0)
    return out


def fcos_assign(geo, boxes, box_labels, box_count, g_max):
    labels = torch.empty((geo.R,), device=boxes.device, dtype=torch.int64)
    call("scan_fcos_assign", geo.ref(), _ptr(boxes), _ptr(box_labels), _ptr(box_count), g_max, _ptr(labels), _stream())
    return labels


def fcos_assign_reg(geo, boxes, box_labels, box_count, g_max):
    """Assignment + (l, t, r, b) regression targets [R,4] for FCOSLossComputation (loss.py:86-126)."""
    labels = torch.empty((geo.R,), device=boxes.device, dtype=torch.int64)
    reg = torch.empty((geo.R, 4), device=boxes.device, dtype=torch.float32)
    call("scan_fcos_assign_reg", geo.ref(), _ptr(boxes), _ptr(box_labels), _ptr(box_count), g_max, _ptr(labels), _ptr(reg), _stream())
    return labels, reg


class _FcosLoss(torch.autograd.Function):
    """(cls_loss, reg_loss, centerness_loss) of FCOSLossComputation.__call__ (loss.py:168-230) in one pass over the head's maps."""

    @staticmethod
    def forward(ctx, geo, labels, reg_targets, gamma, alpha, n_levels, *maps):
        maps = [m.contiguous() for m in maps]
        cls, reg, ctr = maps[:n_levels], maps[n_levels:2 * n_levels], maps[2 * n_levels:]
        dev = cls[0].device
        num_classes = cls[0].shape[1]
        partials = torch.empty((_lib.lib().scan_fcos_loss_num_partials(),), device=dev, dtype=torch.float64)
        sums = torch.empty((6,), device=dev, dtype=torch.float64)
        losses = torch.empty((3,), device=dev, dtype=torch.float32)
        call("scan_fcos_loss_fwd", geo.ref(), _ptr_array(cls), _ptr_array(reg), _ptr_array(ctr), _ptr(labels), _ptr(reg_targets), num_classes,
             gamma, alpha, _ptr(partials), _ptr(sums), _ptr(losses), _stream())
        ctx.geo, ctx.cfg = geo, (gamma, alpha, n_levels, num_classes)
        ctx.save_for_backward(labels, reg_targets, sums, *maps)
        return losses

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_losses):
        labels, reg_targets, sums = ctx.saved_tensors[:3]
        maps = ctx.saved_tensors[3:]
        gamma, alpha, n_levels, num_classes = ctx.cfg
        cls, reg, ctr = maps[:n_levels], maps[n_levels:2 * n_levels], maps[2 * n_levels:]
        d_losses = d_losses.to(torch.float32).contiguous()
        grads = [torch.empty_like(m) for m in maps]
        call("scan_fcos_loss_bwd", ctx.geo.ref(), _ptr_array(cls), _ptr_array(reg), _ptr_array(ctr), _ptr(labels), _ptr(reg_targets),
             num_classes, gamma, alpha, _ptr(sums), _ptr(d_losses), _ptr_array(grads[:n_levels]), _ptr_array(grads[n_levels:2 * n_levels]),
             _ptr_array(grads[2 * n_levels:]), _stream())
        return (None, None, None, None, None, None) + tuple(grads)


def fcos_loss(geo, labels, reg_targets, box_cls, box_regression, centerness, gamma, alpha):
    """Returns a [3] tensor (cls_loss, reg_loss, centerness_loss)."""
    for m in list(box_cls) + list(box_regression) + list(centerness):
        if not m.is_cuda or m.dtype != torch.float32:
            raise RuntimeError("fcos_loss expects CUDA fp32 maps (no CPU fallback)")
    n_levels = len(box_cls)
    return _FcosLoss.apply(geo, labels, reg_targets, float(gamma), float(alpha), n_levels, *box_cls, *box_regression, *centerness)


class SampleResult(object):
    __slots__ = ("node_rows", "node_labels", "meta", "n_nodes")


def sample_nodes(geo, mode, with_bg, labels=None, pos_mask=None, plabel=None):
    """Returns SampleResult (one device->host read of the 176-byte meta record)."""
    dev = (labels if labels is not None else pos_mask).device
    cap = 2 * geo.R
    node_rows = torch.empty((cap,), device=dev, dtype=torch.int32)
    node_labels = torch.empty((cap,), device=dev, dtype=torch.int64)
    meta = torch.zeros((ctypes.sizeof(ScanSampleMeta) // 4,), device=dev, dtype=torch.int32)
    ws_bytes = _lib.lib().scan_sample_workspace_bytes(geo.R)
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    call("scan_sample_nodes", geo.ref(), mode, int(bool(with_bg)), _ptr(labels), _ptr(pos_mask), _ptr(plabel),
         _ptr(node_rows), _ptr(node_labels), cap, _ptr(meta), _ptr(ws), ws_bytes, _stream())
    host = meta.cpu().numpy()
    m = ScanSampleMeta.from_buffer_copy(host.tobytes())
    if m.error == 1:
        raise IndexError("no negative location left at a level with positives (reference: loss.py:503-504)")
    if m.error == 2:
        raise RuntimeError("scan_sample_nodes: node capacity exceeded")
    out = SampleResult()
    out.meta = m
    out.n_nodes = int(m.n_nodes)
    out.node_rows = node_rows[:out.n_nodes]
    out.node_labels = node_labels[:out.n_nodes]
    return out


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, node_rows):
        rows = rows.contiguous()
        m = node_rows.numel()
        out = torch.empty((m, rows.shape[1]), device=rows.device, dtype=torch.float32)
        call("scan_gather_rows", _ptr(rows), _ptr(node_rows), m, rows.shape[1], _ptr(out), _stream())
        ctx.save_for_backward(node_rows)
        ctx.n_rows = rows.shape[0]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_nodes):
        (node_rows,) = ctx.saved_tensors
        d_nodes = d_nodes.contiguous()
        d_rows = torch.zeros((ctx.n_rows, d_nodes.shape[1]), device=d_nodes.device, dtype=torch.float32)
        call("scan_scatter_add_rows", _ptr(d_nodes), _ptr(node_rows), node_rows.numel(), d_nodes.shape[1],
             _ptr(d_rows), _stream())
        return d_rows, None


class _GatherRowsThrough(torch.autograd.Function):
    """gather_rows that also hands `rows` through: (nodes, rows_alias).  Consumers that come AFTER the gather in the forward
    (the conditional convolution of the source branch) read rows_alias, so their dense d_rows arrives HERE together with
    d_nodes and the node gradients are scatter-added straight into it -- instead of a zero-filled [R,256] scatter target plus
    autograd's full-size gradient sum (one 183 MB fill and one 550 MB add per source pass at 8 images)."""

    @staticmethod
    def forward(ctx, rows, node_rows):
        rows = rows.contiguous()
        m = node_rows.numel()
        out = torch.empty((m, rows.shape[1]), device=rows.device, dtype=torch.float32)
        call("scan_gather_rows", _ptr(rows), _ptr(node_rows), m, rows.shape[1], _ptr(out), _stream())
        ctx.save_for_backward(node_rows)
        ctx.n_rows = rows.shape[0]
        ctx.set_materialize_grads(False)
        return out, rows

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_nodes, d_rows):
        (node_rows,) = ctx.saved_tensors
        if d_nodes is None:
            return d_rows, None
        d_nodes = d_nodes.contiguous()
        if d_rows is None:
            d_rows = torch.zeros((ctx.n_rows, d_nodes.shape[1]), device=d_nodes.device, dtype=torch.float32)
        elif not d_rows.is_contiguous():
            d_rows = d_rows.contiguous()
        # d_rows is the fresh buffer the downstream backward just wrote (sole consumer of rows_alias): accumulate in place
        call("scan_scatter_add_rows", _ptr(d_nodes), _ptr(node_rows), node_rows.numel(), d_nodes.shape[1], _ptr(d_rows), _stream())
        return d_rows, None


def gather_rows_through(rows, node_rows):
    """Returns (nodes [M,C], rows_alias): use rows_alias for everything that reads `rows` after this call."""
    return _GatherRowsThrough.apply(rows, node_rows)


def gather_rows(rows, node_rows):
    return _GatherRows.apply(rows, node_rows)


# ----------------------------------------------------------------------------------------------------
# K4b: conditional convolution + activation + focal loss
# ----------------------------------------------------------------------------------------------------
# 0 = tcgen05 (product), 1 = fp32 FFMA verification kernel (tests / bring-up only; SCAN_B200_CONDCONV_IMPL=1)
CONDCONV_IMPL = {"impl": int(__import__("os").environ.get("SCAN_B200_CONDCONV_IMPL", "0"))}


class _CondConv(torch.autograd.Function):
    """through = True: `rows` is also handed on as an alias (last output).  A consumer that runs AFTER the conditional convolution
    (head_out) reads the alias, so its dense d_rows arrives here and this backward adds its own gradient into that buffer inside
    the kernel (scan_condconv_bwd2, accumulate_rows) -- instead of autograd summing two [R,256] gradients with an add kernel."""

    @staticmethod
    def forward(ctx, geo, num_classes, act_mode, loss_weight, rows, weight, bias, labels, through=False):
        rows = rows.contiguous()
        weight = weight.contiguous()
        dev = rows.device
        acts = [torch.empty((geo.n_images, num_classes, h, w), device=dev, dtype=torch.float32) for h, w in geo.shapes]
        n_part = _lib.lib().scan_condconv_num_partials()
        partials = torch.empty((n_part,), device=dev, dtype=torch.float64) if labels is not None else None
        flags = torch.empty((n_part,), device=dev, dtype=torch.int32)
        call("scan_condconv_fwd", geo.ref(), _ptr(rows), _ptr(weight), _ptr(bias), num_classes, act_mode,
             _ptr_array(acts), _ptr(labels), _ptr(partials), _ptr(flags), CONDCONV_IMPL["impl"], _stream())
        norm = float(geo.R) if act_mode == 0 else float(geo.R * num_classes)
        if labels is not None:
            loss = (partials.sum() * (loss_weight / norm)).to(torch.float32)
        else:
            loss = torch.zeros((), device=dev, dtype=torch.float32)
        ctx.geo, ctx.k, ctx.act_mode = geo, num_classes, act_mode
        ctx.loss_scale = loss_weight / norm if labels is not None else 0.0
        ctx.has_bias = bias is not None
        ctx.save_for_backward(rows, weight, labels, *acts)
        ctx.mark_non_differentiable(flags)
        ctx.through = bool(through)
        if through:
            ctx.set_materialize_grads(False)
            return (loss, flags) + tuple(acts) + (rows,)
        return (loss, flags) + tuple(acts)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_loss, _d_flags, *d_acts):
        rows, weight, labels = ctx.saved_tensors[:3]
        acts = ctx.saved_tensors[3:]
        geo, k = ctx.geo, ctx.k
        dev = rows.device
        d_alias = None
        if ctx.through:
            d_alias, d_acts = d_acts[-1], d_acts[:-1]
            if d_loss is None and all(g is None for g in d_acts):      # nothing reached the maps or the loss: hand the alias' gradient on
                return None, None, None, None, d_alias, None, None, None, None
        d_acts = [None if g is None else g.contiguous() for g in d_acts]
        # d(total)/d(act_loss) stays on the device (read by the kernel): no host sync in backward
        d_loss_dev = None
        if ctx.loss_scale != 0.0 and d_loss is not None:
            d_loss_dev = d_loss.to(torch.float32).reshape(1).contiguous()
        # d_alias is the fresh buffer the later consumer's backward just wrote (sole consumer of the alias): accumulate in place
        acc = d_alias is not None
        d_rows = (d_alias if d_alias.is_contiguous() else d_alias.contiguous()) if acc else torch.empty_like(rows)
        d_weight = torch.empty_like(weight)
        d_bias = torch.empty((k,), device=dev, dtype=torch.float32) if ctx.has_bias else None
        ws_bytes = _lib.lib().scan_condconv_bwd_workspace_bytes(geo.ref(), k)
        ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
        call("scan_condconv_bwd2", geo.ref(), _ptr(rows), _ptr(weight), k, ctx.act_mode, _ptr_array(acts),
             _ptr_array(d_acts), _ptr(labels), ctx.loss_scale, _ptr(d_loss_dev), _ptr(d_rows), int(acc), _ptr(d_weight),
             _ptr(d_bias), _ptr(ws), ws_bytes, _stream())
        return None, None, None, None, d_rows, d_weight, d_bias, None, None


def condconv(geo, rows, weight, bias, num_classes, act_mode, labels=None, loss_weight=1.0, through=False):
    """Returns (act_maps: list of [N,K,H_l,W_l], loss or None, flags) and, with through=True, an alias of `rows` that every LATER
    consumer of the rows must read (see _CondConv)."""
    if num_classes > _lib.SCAN_MAX_CLASSES:
        raise RuntimeError("used_num_classes > %d is not supported" % _lib.SCAN_MAX_CLASSES)
    out = _CondConv.apply(geo, num_classes, act_mode, float(loss_weight), rows, weight, bias, labels, bool(through))
    if through:
        loss, flags, acts, alias = out[0], out[1], list(out[2:-1]), out[-1]
        return acts, (loss if labels is not None else None), flags, alias
    loss, flags, acts = out[0], out[1], list(out[2:])
    return acts, (loss if labels is not None else None), flags


# ----------------------------------------------------------------------------------------------------
# K4a: manifestation (RNN variant)
# ----------------------------------------------------------------------------------------------------
class _ManifestRnn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proto, w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, wc, bc):
        ps = [t.contiguous() for t in (w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, wc, bc)]
        proto = proto.contiguous()
        k, i, p = proto.shape
        h, o = ps[0].shape[0], ps[8].shape[0]
        dev = proto.device
        out = torch.empty((k, o), device=dev, dtype=torch.float32)
        saved = torch.empty((_lib.lib().scan_manifest_rnn_saved_floats(k, p, i, h),), device=dev, dtype=torch.float32)
        call("scan_manifest_rnn_fwd", _ptr(proto), k, p, i, h, o, *[_ptr(t) for t in ps], _ptr(out), _ptr(saved), _stream())
        ctx.dims = (k, p, i, h, o)
        ctx.save_for_backward(saved, *ps)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_out):
        k, p, i, h, o = ctx.dims
        saved = ctx.saved_tensors[0]
        w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, wc, bc = ctx.saved_tensors[1:]
        d_out = d_out.contiguous()
        grads = [torch.empty_like(t) for t in (w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, wc, bc)]
        ws = torch.empty((_lib.lib().scan_manifest_rnn_workspace_bytes(k, p, h, o),), device=d_out.device, dtype=torch.uint8)
        call("scan_manifest_rnn_bwd", _ptr(d_out), k, p, i, h, o, _ptr(w_hh0), _ptr(w_ih1), _ptr(w_hh1), _ptr(wc), _ptr(saved),
             *[_ptr(g) for g in grads], _ptr(ws), ws.numel(), _stream())
        return (None,) + tuple(grads)


def manifest_rnn(proto, rnn, conv):
    """get_conded_weight (condgraph.py:313-336), RNN variant: `rnn` = nn.RNN(I, H, 2, tanh), `conv` = cond_nx1 (P x 1)."""
    if rnn.num_layers != 2 or rnn.nonlinearity != "tanh" or rnn.bidirectional or rnn.batch_first or not rnn.bias:
        raise RuntimeError("manifest_rnn implements the reference's nn.RNN(256, 512, 2, nonlinearity='tanh')")
    if not proto.is_cuda or proto.dtype != torch.float32:
        raise RuntimeError("manifest_rnn expects a CUDA fp32 prototype buffer (no CPU fallback)")
    return _ManifestRnn.apply(proto, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0, rnn.weight_ih_l1,
                              rnn.weight_hh_l1, rnn.bias_ih_l1, rnn.bias_hh_l1, conv.weight, conv.bias)


# ----------------------------------------------------------------------------------------------------
# K4a': manifestation without the RNN (tiny-batch dense layers over the K paradigm rows, csrc/rowsmlp.cu)
# ----------------------------------------------------------------------------------------------------
class _RowsLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu):
        x, w = x.contiguous(), w.contiguous()
        b = None if b is None else b.contiguous()
        k, i = x.shape
        o = w.shape[0]
        y = torch.empty((k, o), device=x.device, dtype=torch.float32)
        call("scan_rows_linear_fwd", _ptr(x), _ptr(w), _ptr(b), k, i, o, int(relu), _ptr(y), _stream())
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (int(relu), b is not None, x.requires_grad)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        relu, has_b, need_dx = ctx.cfg
        dy = dy.contiguous()
        k, i = x.shape
        o = w.shape[0]
        d_w = torch.empty_like(w)
        d_b = torch.empty((o,), device=x.device, dtype=torch.float32) if has_b else None
        d_x = torch.empty_like(x) if need_dx else None
        call("scan_rows_linear_bwd", _ptr(x), _ptr(w), _ptr(dy), _ptr(y), k, i, o, relu, _ptr(d_w), _ptr(d_b), _ptr(d_x), _stream())
        return d_x, d_w, d_b, None


def rows_linear(x, weight, bias, relu=False):
    """act(x @ weight.T + bias) for a tiny batch x [K<=16, I] (weight may be a conv weight [O, I', P, 1] flattened to [O, I' P])."""
    if not x.is_cuda or x.dtype != torch.float32 or x.shape[0] > _lib.SCAN_MAX_CLASSES:
        raise RuntimeError("rows_linear expects a CUDA fp32 [K<=16, I] input (no CPU fallback)")
    return _RowsLinear.apply(x, weight.reshape(weight.shape[0], -1), bias, relu)


class _RowsGnRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps):
        x, gamma, beta = x.contiguous(), gamma.contiguous(), beta.contiguous()
        k, c = x.shape
        y = torch.empty_like(x)
        stats = torch.empty((k, groups, 2), device=x.device, dtype=torch.float32)
        call("scan_rows_gn_relu_fwd", _ptr(x), _ptr(gamma), _ptr(beta), k, c, groups, eps, _ptr(y), _ptr(stats), _stream())
        ctx.save_for_backward(x, y, gamma, stats)
        ctx.groups = groups
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, y, gamma, stats = ctx.saved_tensors
        dy = dy.contiguous()
        k, c = x.shape
        d_x, d_g, d_b = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(gamma)
        call("scan_rows_gn_relu_bwd", _ptr(x), _ptr(y), _ptr(dy), _ptr(gamma), _ptr(stats), k, c, ctx.groups, _ptr(d_x), _ptr(d_g), _ptr(d_b),
             _stream())
        return d_x, d_g, d_b, None, None


def rows_gn_relu(x, gamma, beta, groups, eps):
    return _RowsGnRelu.apply(x, gamma, beta, int(groups), float(eps))


# ----------------------------------------------------------------------------------------------------
# K3a: attention
# ----------------------------------------------------------------------------------------------------
# implementations: "t5" = tcgen05 kernels (product), "ffma" = fp32 verification kernels (SCAN_B200_ATTN_FWD/BWD=ffma)
ATTN_IMPL = {"fwd": __import__("os").environ.get("SCAN_B200_ATTN_FWD", "t5"),
             "bwd": __import__("os").environ.get("SCAN_B200_ATTN_BWD", "t5")}


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, scale, drop_p, seed):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        m = q.shape[0]
        out = torch.empty_like(q)
        lse = torch.empty((4 * m,), device=q.device, dtype=torch.float32)
        ws = None
        if ATTN_IMPL["fwd"] == "t5":
            ws = torch.empty((_lib.lib().scan_attn_workspace_bytes(m),), device=q.device, dtype=torch.uint8)
        call("scan_attn_fwd", _ptr(q), _ptr(k), _ptr(v), m, scale, drop_p, seed, _ptr(out), _ptr(lse), _ptr(ws),
             0 if ws is None else ws.numel(), _stream())
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.cfg = (scale, drop_p, seed)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_out):
        q, k, v, out, lse = ctx.saved_tensors
        scale, drop_p, seed = ctx.cfg
        d_out = d_out.contiguous()
        m = q.shape[0]
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        delta = torch.empty((4 * m,), device=q.device, dtype=torch.float32)
        ws = None
        if ATTN_IMPL["bwd"] == "t5":
            ws = torch.empty((_lib.lib().scan_attn_bwd_workspace_bytes(m),), device=q.device, dtype=torch.uint8)
        call("scan_attn_bwd", _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse), _ptr(d_out), m, scale, drop_p, seed,
             _ptr(dq), _ptr(dk), _ptr(dv), _ptr(delta), _ptr(ws), 0 if ws is None else ws.numel(), _stream())
        return dq, dk, dv, None, None, None


def chunked_attention(q, k, v, scale=0.25, drop_p=0.0, seed=0):
    if q.shape[1] != C:
        raise RuntimeError("attention expects [M,256] projections (4 chunks of 64-d sub-tokens)")
    return _Attention.apply(q, k, v, float(scale), float(drop_p), int(seed))


# ----------------------------------------------------------------------------------------------------
# K3a': the dense layers around the attention core and the node classifier (tcgen05 3xTF32 GEMMs, csrc/gemm.cu)
# ----------------------------------------------------------------------------------------------------
def _graph_ws(m, device):
    return torch.empty((_lib.lib().scan_graph_workspace_bytes(m),), device=device, dtype=torch.uint8)


class _GraphAttention(torch.autograd.Function):
    """MultiHeadAttention.forward (layers/transformer.py:53-90) for key = value = query = x [M,256]:
    q,k,v projections -> chunked attention -> linear_final -> dropout -> LayerNorm(x + .), one autograd node."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, wf, bf, gamma, beta, scale, drop_p, seed, eps):
        x = x.contiguous()
        m = x.shape[0]
        dev = x.device
        w_qkv = torch.cat([wq, wk, wv]).contiguous()            # [768,256]: q | k | v output parts
        b_qkv = torch.cat([bq, bk, bv]).contiguous()
        wf, bf, gamma, beta = wf.contiguous(), bf.contiguous(), gamma.contiguous(), beta.contiguous()
        qkv = torch.empty((3, m, C), device=dev, dtype=torch.float32)
        call("scan_qkv_fwd", _ptr(x), _ptr(w_qkv), _ptr(b_qkv), m, _ptr(qkv), _stream())
        att = torch.empty((m, C), device=dev, dtype=torch.float32)
        lse = torch.empty((4 * m,), device=dev, dtype=torch.float32)
        ws = torch.empty((_lib.lib().scan_attn_workspace_bytes(m),), device=dev, dtype=torch.uint8)
        call("scan_attn_fwd", _ptr(qkv[0]), _ptr(qkv[1]), _ptr(qkv[2]), m, scale, drop_p, seed, _ptr(att), _ptr(lse), _ptr(ws),
             ws.numel(), _stream())
        y = torch.empty((m, C), device=dev, dtype=torch.float32)
        xhat = torch.empty((m, C), device=dev, dtype=torch.float32)
        rstd = torch.empty((m,), device=dev, dtype=torch.float32)
        call("scan_attn_out_ln_fwd", _ptr(att), _ptr(wf), _ptr(bf), _ptr(x), _ptr(gamma), _ptr(beta), m, eps, drop_p,
             seed ^ 0x5DEECE66D, _ptr(y), _ptr(xhat), _ptr(rstd), _stream())
        ctx.save_for_backward(x, w_qkv, qkv, att, lse, wf, gamma, xhat, rstd)
        ctx.cfg = (scale, drop_p, seed)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_y):
        x, w_qkv, qkv, att, lse, wf, gamma, xhat, rstd = ctx.saved_tensors
        scale, drop_p, seed = ctx.cfg
        m = x.shape[0]
        dev = x.device
        d_y = d_y.contiguous()
        ws = _graph_ws(m, dev)
        d_x = torch.empty_like(x)
        d_att = torch.empty_like(x)
        d_wf = torch.empty_like(wf)
        d_bf = torch.empty((C,), device=dev, dtype=torch.float32)
        d_gb = torch.empty((2 * C,), device=dev, dtype=torch.float32)
        call("scan_attn_out_ln_bwd", _ptr(d_y), _ptr(xhat), _ptr(rstd), _ptr(gamma), _ptr(att), _ptr(wf), m, drop_p,
             seed ^ 0x5DEECE66D, _ptr(d_x), _ptr(d_att), _ptr(d_wf), _ptr(d_bf), _ptr(d_gb), _ptr(ws), ws.numel(), _stream())
        d_qkv = torch.empty_like(qkv)
        delta = torch.empty((4 * m,), device=dev, dtype=torch.float32)
        ws2 = torch.empty((_lib.lib().scan_attn_bwd_workspace_bytes(m),), device=dev, dtype=torch.uint8)
        call("scan_attn_bwd", _ptr(qkv[0]), _ptr(qkv[1]), _ptr(qkv[2]), _ptr(att), _ptr(lse), _ptr(d_att), m, scale, drop_p, seed,
             _ptr(d_qkv[0]), _ptr(d_qkv[1]), _ptr(d_qkv[2]), _ptr(delta), _ptr(ws2), ws2.numel(), _stream())
        d_w = torch.empty_like(w_qkv)
        d_b = torch.empty((3 * C,), device=dev, dtype=torch.float32)
        call("scan_qkv_bwd", _ptr(d_qkv), _ptr(x), _ptr(w_qkv), m, 1, _ptr(d_x), _ptr(d_w), _ptr(d_b), _ptr(ws), ws.numel(), _stream())
        return (d_x, d_w[:C], d_b[:C], d_w[C:2 * C], d_b[C:2 * C], d_w[2 * C:], d_b[2 * C:], d_wf, d_bf, d_gb[:C], d_gb[C:],
                None, None, None, None)


def graph_attention(x, attn, drop_p, seed):
    """attn: module holding linear_q/k/v/final + layer_norm (layers/transformer.py:36-52)."""
    if x.shape[1] != C or not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("graph_attention expects CUDA fp32 [M,256] nodes (no CPU fallback)")
    scale = float((attn.dim_per_head // attn.num_heads) ** -0.5)          # transformer.py:75 -> 0.25
    return _GraphAttention.apply(x, attn.linear_q.weight, attn.linear_q.bias, attn.linear_k.weight, attn.linear_k.bias,
                                 attn.linear_v.weight, attn.linear_v.bias, attn.linear_final.weight, attn.linear_final.bias,
                                 attn.layer_norm.weight, attn.layer_norm.bias, scale, float(drop_p), int(seed),
                                 float(attn.layer_norm.eps))


class _NodeClassifier(torch.autograd.Function):
    """loss_weight * CE(proto_cls(relu(proto_cls_hidden(nodes))), labels - shift) (condgraph.py:400-402)."""

    @staticmethod
    def forward(ctx, nodes, w1, b1, w2, b2, labels, shift, loss_weight):
        nodes = nodes.contiguous()
        w1, b1, w2, b2 = w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous()
        m, dev = nodes.shape[0], nodes.device
        k, h = w2.shape[0], w1.shape[0]
        hidden = torch.empty((m, h), device=dev, dtype=torch.float32)
        dlogits = torch.empty((m, 16), device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        ws = torch.empty((1024 + 8 * ((m + 127) // 128),), device=dev, dtype=torch.uint8)
        call("scan_node_cls_fwd", _ptr(nodes), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(labels), m, h, k, shift, loss_weight,
             _ptr(hidden), _ptr(dlogits), _ptr(loss), _ptr(ws), ws.numel(), _stream())
        ctx.save_for_backward(nodes, w1, w2, hidden, dlogits)
        ctx.cfg = (k, h, loss_weight)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_loss):
        nodes, w1, w2, hidden, dlogits = ctx.saved_tensors
        k, h, loss_weight = ctx.cfg
        m, dev = nodes.shape[0], nodes.device
        d_loss = d_loss.to(torch.float32).reshape(1).contiguous()
        ws = _graph_ws(m, dev)
        d_nodes = torch.empty_like(nodes)
        d_w1, d_w2 = torch.empty_like(w1), torch.empty_like(w2)
        d_b1 = torch.empty((h,), device=dev, dtype=torch.float32)
        d_b2 = torch.empty((k,), device=dev, dtype=torch.float32)
        call("scan_node_cls_bwd", _ptr(dlogits), _ptr(hidden), _ptr(nodes), _ptr(w1), _ptr(w2), m, h, k, loss_weight, _ptr(d_loss),
             _ptr(d_nodes), _ptr(d_w1), _ptr(d_b1), _ptr(d_w2), _ptr(d_b2), _ptr(ws), ws.numel(), _stream())
        return d_nodes, d_w1, d_b1, d_w2, d_b2, None, None, None


def node_classifier_loss(nodes, hidden_layer, out_layer, labels, label_shift, loss_weight):
    if nodes.shape[1] != C or hidden_layer.weight.shape != (512, C) or out_layer.weight.shape[0] > _lib.SCAN_MAX_CLASSES:
        raise RuntimeError("node classifier kernel is built for Linear(256,512) -> Linear(512,K<=16)")
    if not nodes.is_cuda:
        raise RuntimeError("node_classifier_loss needs CUDA tensors (no CPU fallback)")
    if nodes.shape[0] == 0:
        raise RuntimeError("node_classifier_loss: no nodes")
    return _NodeClassifier.apply(nodes, hidden_layer.weight, hidden_layer.bias, out_layer.weight, out_layer.bias, labels,
                                 int(label_shift), float(loss_weight))


class _ClassMeans(torch.autograd.Function):
    """prototype_batch[c] = nodes[labels == c + shift].mean(0), zero for absent classes (condgraph.py:395-398), differentiable."""

    @staticmethod
    def forward(ctx, nodes, labels, num_classes, shift):
        nodes = nodes.contiguous()
        packed = class_sums(nodes, labels, num_classes, shift)
        cnt = packed[:, -1:]
        means = torch.where(cnt > 0, packed[:, :-1] / cnt.clamp(min=1.0), torch.zeros_like(packed[:, :-1]))
        ctx.save_for_backward(packed, labels)
        ctx.cfg = (nodes.shape[0], nodes.shape[1], shift)
        ctx.mark_non_differentiable(packed)
        return means, packed

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_means, _d_packed):
        packed, labels = ctx.saved_tensors
        m, c, shift = ctx.cfg
        d_nodes = torch.empty((m, c), device=packed.device, dtype=torch.float32)
        d_means = d_means.contiguous()
        call("scan_class_mean_bwd", _ptr(d_means), _ptr(packed), _ptr(labels), m, c, shift, _ptr(d_nodes), _stream())
        return d_nodes, None, None, None


def class_means(nodes, labels, num_classes, label_shift):
    """Returns (means [K,C] with gradient, packed [K,C+1] sum|count buffer)."""
    return _ClassMeans.apply(nodes, labels, int(num_classes), int(label_shift))


# ----------------------------------------------------------------------------------------------------
# a6: per-class GCN (GLOBAL_GCN = False)
# ----------------------------------------------------------------------------------------------------
def _pad(v, q):
    return (v + q - 1) // q * q


def gemm_nt(a, b, m, n, k, lda, ldb, out=None, ldc=None, bias=None, relu=False, accumulate=False):
    """out [m, n] (+)= act(a [m, k] @ b [n, k].T + bias) on the tcgen05 3xTF32 GEMM (raw-pointer form: pitches in floats)."""
    dev = a.device
    if out is None:
        ldc = _pad(n, 4)
        out = torch.empty((m, ldc), device=dev, dtype=torch.float32)
    ws_bytes = _lib.lib().scan_gemm_nt_workspace_bytes(m, n, k)
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    call("scan_gemm_nt", _ptr(a), lda, _ptr(b), ldb, m, n, k, _ptr(bias), int(relu), int(accumulate), _ptr(out), ldc, _ptr(ws), ws_bytes,
         _stream())
    return out


def transpose_pad(x, n_rows, n_cols, ld_src, ld_dst):
    out = torch.empty((n_cols, ld_dst), device=x.device, dtype=torch.float32)
    call("scan_transpose", _ptr(x), n_rows, n_cols, ld_src, _ptr(out), ld_dst, _stream())
    return out


GCN_ACT = {"NO": 0, "relu": 1, "sigmoid": 2, "tanh": 3, "softmax": 4}


class _LocalGCN(torch.autograd.Function):
    """The per-class graph convolution of condgraph.py:404-414 with GCNs (:262-282) and get_edge (:284-302):
    for every class c with members idx_c:  Adj = softmax(affinity(X_c, X_c)).detach();  X1 = relu(W1 (Adj X_c) + b1);
    Y = act(W2 (Adj X1) + b2) (+ X_c);  out[idx_c] = Y.  All products run on the tcgen05 3xTF32 GEMM (scan_gemm_nt), the
    class blocks are gathered / scattered with the node kernels; gradients reach the nodes only through the second operand
    of Adj . (the adjacency is detached in the reference)."""

    @staticmethod
    def forward(ctx, points, w1, b1, w2, b2, cosine, act_mode, shortcut, idx_list):
        points = points.contiguous()
        w1, b1, w2, b2 = w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous()
        dev = points.device
        out = torch.zeros_like(points)
        saved = []
        for idx in idx_list:
            mc = idx.numel()
            mp4, mp32 = _pad(mc, 4), _pad(mc, 32)
            sub = torch.empty((mc, C), device=dev, dtype=torch.float32)
            call("scan_gather_rows", _ptr(points), _ptr(idx), mc, C, _ptr(sub), _stream())
            base = sub
            if cosine:
                base = torch.empty_like(sub)
                call("scan_rows_l2normalize", _ptr(sub), mc, 1e-8, _ptr(base), _stream())
            adj = gemm_nt(base, base, mc, mc, C, C, C)                       # [mc, mp4] affinity
            call("scan_rows_softmax", _ptr(adj), mc, mc, mp4, _stream())
            sub_t = transpose_pad(sub, mc, C, C, mp32)                       # [256, mp32]
            h1 = gemm_nt(adj, sub_t, mc, C, mc, mp4, mp32)                   # Adj . X
            x1 = gemm_nt(h1, w1, mc, C, C, C, C, bias=b1, relu=True)
            x1_t = transpose_pad(x1, mc, C, C, mp32)
            h2 = gemm_nt(adj, x1_t, mc, C, mc, mp4, mp32)
            z2 = gemm_nt(h2, w2, mc, C, C, C, C, bias=b2)
            act = torch.empty_like(sub)
            y = torch.empty_like(sub) if shortcut else act
            call("scan_gcn_act_fwd", _ptr(z2), _ptr(sub) if shortcut else None, mc, act_mode, _ptr(act), _ptr(y), _stream())
            call("scan_scatter_add_rows", _ptr(y), _ptr(idx), mc, C, _ptr(out), _stream())
            saved.append((idx, adj, h1, x1, h2, act))
        ctx.saved = saved
        ctx.cfg = (act_mode, shortcut)
        ctx.save_for_backward(w1, w2)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_out):
        w1, w2 = ctx.saved_tensors
        act_mode, shortcut = ctx.cfg
        d_out = d_out.contiguous()
        dev = d_out.device
        d_points = torch.zeros_like(d_out)
        d_w1, d_w2 = torch.zeros_like(w1), torch.zeros_like(w2)
        d_b1 = torch.zeros((w1.shape[0],), device=dev, dtype=torch.float32)
        d_b2 = torch.zeros((w2.shape[0],), device=dev, dtype=torch.float32)
        w1_t = transpose_pad(w1, C, C, C, C)
        w2_t = transpose_pad(w2, C, C, C, C)
        for idx, adj, h1, x1, h2, act in ctx.saved:
            mc = idx.numel()
            mp4, mp32 = _pad(mc, 4), _pad(mc, 32)
            dy = torch.empty((mc, C), device=dev, dtype=torch.float32)
            call("scan_gather_rows", _ptr(d_out), _ptr(idx), mc, C, _ptr(dy), _stream())
            dz2 = torch.empty_like(dy)
            call("scan_gcn_act_bwd", _ptr(act), _ptr(dy), mc, act_mode, _ptr(dz2), _stream())
            ws_bytes = _lib.lib().scan_linear_wgrad_workspace_bytes(mc, C, C)
            ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
            call("scan_linear_wgrad", _ptr(dz2), _ptr(h2), mc, C, C, 1, _ptr(d_w2), _ptr(d_b2), _ptr(ws), ws_bytes, _stream())
            dh2 = gemm_nt(dz2, w2_t, mc, C, C, C, C)
            adj_t = transpose_pad(adj, mc, mc, mp4, mp4)
            dh2_t = transpose_pad(dh2, mc, C, C, mp32)
            dx1 = gemm_nt(adj_t, dh2_t, mc, C, mc, mp4, mp32)
            dz1 = torch.empty_like(dy)
            call("scan_gcn_act_bwd", _ptr(x1), _ptr(dx1), mc, 1, _ptr(dz1), _stream())
            call("scan_linear_wgrad", _ptr(dz1), _ptr(h1), mc, C, C, 1, _ptr(d_w1), _ptr(d_b1), _ptr(ws), ws_bytes, _stream())
            dh1 = gemm_nt(dz1, w1_t, mc, C, C, C, C)
            dh1_t = transpose_pad(dh1, mc, C, C, mp32)
            d_sub = dy if shortcut else torch.zeros_like(dy)
            gemm_nt(adj_t, dh1_t, mc, C, mc, mp4, mp32, out=d_sub, ldc=C, accumulate=True)
            call("scan_scatter_add_rows", _ptr(d_sub), _ptr(idx), mc, C, _ptr(d_points), _stream())
        return d_points, d_w1, d_b1, d_w2, d_b2, None, None, None, None


def local_gcn(points, labels, layer1, layer2, num_classes, label_shift, edge_norm, out_act, shortcut):
    """Per-class GCN over the sampled nodes.  The class membership lists are built on the host (one D2H of the label vector;
    the reference synchronises once per class at `.any()`, condgraph.py:406-407)."""
    if points.shape[1] != C or layer1.weight.shape != (C, C) or layer2.weight.shape != (C, C):
        raise RuntimeError("local_gcn is built for 256 -> 256 -> 256 graph convolutions")
    if not points.is_cuda:
        raise RuntimeError("local_gcn needs CUDA tensors (no CPU fallback)")
    if out_act not in GCN_ACT:
        raise KeyError("unknown gcn output activation")
    host = labels.cpu()
    idx_list = []
    for c in range(num_classes):
        idx = torch.nonzero(host == c + label_shift).reshape(-1)
        if idx.numel():
            idx_list.append(idx.to(torch.int32).to(points.device, non_blocking=True))
    return _LocalGCN.apply(points, layer1.weight, layer1.bias, layer2.weight, layer2.bias, edge_norm != "NO", GCN_ACT[out_act],
                           bool(shortcut), idx_list)


# ----------------------------------------------------------------------------------------------------
# a14: transfer losses
# ----------------------------------------------------------------------------------------------------
TRANSFER_FLAGS = {"PROTOTYPE": 1, "ADJ": 2, "ADJ_COMPLETE": 4}


class _TransferLoss(torch.autograd.Function):
    """get_transfer_loss (condgraph.py:457-498): sum of the enabled NODES / PROTOTYPE / ADJ / ADJ_COMPLETE losses as one
    autograd node (three launches forward); gradients flow to the target nodes and to the target class means."""

    @staticmethod
    def forward(ctx, nodes, labels, tg_proto, prototype, with_nodes, proto_flags):
        dev = prototype.device
        k, c = prototype.shape[0], prototype.shape[1]
        p_iter = prototype.shape[2] if prototype.dim() == 3 else 1
        prototype = prototype.contiguous()
        loss_n = diff = None
        if with_nodes:
            nodes = nodes.contiguous()
            m = nodes.shape[0]
            diff = torch.empty_like(nodes)
            partials = torch.empty((_lib.lib().scan_transfer_nodes_num_partials(),), device=dev, dtype=torch.float64)
            loss_n = torch.empty((), device=dev, dtype=torch.float32)
            call("scan_transfer_nodes_fwd", _ptr(nodes), _ptr(labels), _ptr(prototype), p_iter, m, k, _ptr(diff), _ptr(partials),
                 _ptr(loss_n), _stream())
        d_tg = None
        out = loss_n
        if proto_flags:
            tg = tg_proto.contiguous()
            losses = torch.empty((4,), device=dev, dtype=torch.float32)
            d_tg = torch.empty((k, c), device=dev, dtype=torch.float32)
            call("scan_transfer_proto", _ptr(tg), _ptr(prototype), p_iter, k, proto_flags, _ptr(loss_n), _ptr(losses), _ptr(d_tg), _stream())
            out = losses[3]
        ctx.save_for_backward(diff, d_tg)
        ctx.m = nodes.shape[0] if with_nodes else 0
        return out.clone() if proto_flags else out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_loss):
        diff, d_tg = ctx.saved_tensors
        d_loss = d_loss.to(torch.float32).reshape(1).contiguous()
        d_nodes = None
        if diff is not None:
            d_nodes = torch.empty_like(diff)
            call("scan_transfer_nodes_bwd", _ptr(diff), _ptr(d_loss), ctx.m, _ptr(d_nodes), _stream())
        d_tgp = d_tg * d_loss if d_tg is not None else None
        return d_nodes, None, d_tgp, None, None, None


def transfer_loss(cfg_names, nodes, labels, tg_proto, prototype):
    """cfg_names: MODEL.MIDDLE_HEAD.TRANSFER_CFG.  Returns the summed loss tensor or None when no term is enabled."""
    with_nodes = ("NODES" in cfg_names) or ("NODE" in cfg_names)
    flags = sum(v for n, v in TRANSFER_FLAGS.items() if n in cfg_names)
    if not with_nodes and not flags:
        return None
    if prototype.shape[1] != C or prototype.shape[0] > _lib.SCAN_MAX_CLASSES or not prototype.is_cuda:
        raise RuntimeError("transfer_loss expects a CUDA [K<=16, 256(, P)] prototype buffer (no CPU fallback)")
    return _TransferLoss.apply(nodes, labels, tg_proto, prototype, with_nodes, flags)


# ----------------------------------------------------------------------------------------------------
# K3b: prototype sums + EMA
# ----------------------------------------------------------------------------------------------------
def class_sums(nodes, labels, num_classes, label_shift):
    nodes = nodes.detach().contiguous()
    packed = torch.empty((num_classes, nodes.shape[1] + 1), device=nodes.device, dtype=torch.float32)
    call("scan_class_sums", _ptr(nodes), _ptr(labels), nodes.shape[0], nodes.shape[1], num_classes, label_shift,
         _ptr(packed), _stream())
    return packed


def proto_update(packed, prototype, slot, shift, cosine_on, momentum):
    k = prototype.shape[0]
    c = prototype.shape[1]
    p = prototype.shape[2] if prototype.dim() == 3 else 1
    if not prototype.is_contiguous():
        raise RuntimeError("prototype buffer must be contiguous")
    batch = torch.empty((k, c), device=prototype.device, dtype=torch.float32)
    call("scan_proto_update", _ptr(packed), k, c, p, slot, int(shift), int(bool(cosine_on)), momentum,
         _ptr(prototype), _ptr(batch), _stream())
    return batch


# ----------------------------------------------------------------------------------------------------
# K2: DBSCAN
# ----------------------------------------------------------------------------------------------------
def dbscan_workspace(cap, device):
    nbytes = _lib.lib().scan_dbscan_workspace_bytes(cap)
    return torch.empty((nbytes,), device=device, dtype=torch.uint8)


def dbscan_level(rows_level, act, thr, eps, cap, pos_mask, plabel, workspace, min_samples=5):
    """act [N,K,H,W]; writes pos_mask [N*H*W] uint8 / plabel [N*H*W] int64 views; returns (labels int32 [cap], info int32[8])."""
    n, k, h, w = act.shape
    labels = torch.empty((cap,), device=act.device, dtype=torch.int32)
    info = torch.empty((8,), device=act.device, dtype=torch.int32)
    call("scan_dbscan_level", _ptr(rows_level), _ptr(act), n, k, h, w, float(thr), float(eps), min_samples, cap,
         _ptr(pos_mask), _ptr(plabel), _ptr(labels), _ptr(info), _ptr(workspace), workspace.numel(), _stream())
    return labels, info


def dbscan_points(points, eps, min_samples=5):
    points = points.contiguous()
    n, dim = points.shape
    labels = torch.empty((max(n, 1),), device=points.device, dtype=torch.int32)
    info = torch.empty((8,), device=points.device, dtype=torch.int32)
    ws = dbscan_workspace(max(n, 1), points.device)
    call("scan_dbscan_points", _ptr(points), n, dim, float(eps), min_samples, _ptr(labels), _ptr(info), _ptr(ws),
         ws.numel(), _stream())
    return labels[:n], info


# ----------------------------------------------------------------------------------------------------
# K5: sigmoid focal loss, ensembling
# ----------------------------------------------------------------------------------------------------
class _SigmoidFocal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, gamma, alpha):
        logits = logits.contiguous()
        targets = targets.contiguous().to(torch.int32)
        losses = torch.empty_like(logits)
        call("scan_sigmoid_focal_fwd", _ptr(logits), _ptr(targets), logits.shape[0], logits.shape[1], gamma, alpha,
             _ptr(losses), _stream())
        ctx.save_for_backward(logits, targets)
        ctx.cfg = (gamma, alpha)
        return losses

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_losses):
        logits, targets = ctx.saved_tensors
        gamma, alpha = ctx.cfg
        d_losses = d_losses.contiguous()
        d_logits = torch.empty_like(logits)
        call("scan_sigmoid_focal_bwd", _ptr(logits), _ptr(targets), _ptr(d_losses), logits.shape[0], logits.shape[1],
             gamma, alpha, _ptr(d_logits), _stream())
        return d_logits, None, None, None


def sigmoid_focal_loss(logits, targets, gamma, alpha):
    """Element-wise losses [R, C]; mirrors `_C.sigmoid_focalloss_forward/backward` (layers/sigmoid_focal_loss.py:9-37)."""
    if logits.dim() != 2:
        raise RuntimeError("logits must be [R, num_classes]")
    return _SigmoidFocal.apply(logits, targets, float(gamma), float(alpha))


_MODES = {"common": 0, "light": 1, "precision": 2}


def ensemble_levels(mode, cls_logits, acts):
    """TEST.MODE map ensembling of all levels (fcos.py:162-169 + inference.py:68): list of class-probability maps
    [N,K-1,H_l,W_l].  One launch for every level; 'light' is a zero-copy view of the activation maps."""
    if mode == "light":
        return [a[:, 1:] for a in acts]
    acts = [a.contiguous() for a in acts]
    cls = [c.contiguous() for c in cls_logits]
    n, k = acts[0].shape[:2]
    geo = Geometry([tuple(a.shape[-2:]) for a in acts], [1] * len(acts), n)
    outs = [torch.empty((n, k - 1) + tuple(a.shape[-2:]), device=a.device, dtype=torch.float32) for a in acts]
    call("scan_ensemble_levels", geo.ref(), _ptr_array(cls), _ptr_array(acts), k, _MODES[mode], _ptr_array(outs), _stream())
    return outs


def ensemble_levels_common(cls_logits):
    """sigmoid of the classification logits of all levels in one launch (TEST.MODE 'common', inference.py:68)."""
    cls = [c.contiguous() for c in cls_logits]
    n, c = cls[0].shape[:2]
    geo = Geometry([tuple(t.shape[-2:]) for t in cls], [1] * len(cls), n)
    outs = [torch.empty_like(t) for t in cls]
    call("scan_ensemble_levels", geo.ref(), _ptr_array(cls), None, c + 1, 0, _ptr_array(outs), _stream())
    return outs


def ensemble(mode, cls_logits, act):
    """One level (kept for callers that hold a single map)."""
    return ensemble_levels(mode, None if cls_logits is None else [cls_logits], [act])[0]


# ----------------------------------------------------------------------------------------------------
# f4: FCOS post-processor
# ----------------------------------------------------------------------------------------------------
PP_CAP = 8192


def postprocess(geo, probs, box_regression, centerness, image_sizes, num_classes_fg, pre_nms_thresh, pre_nms_top_n, nms_thresh,
                post_top_n, min_size=0.0):
    """FCOSPostProcessor.forward (inference.py:54-194) on per-level class PROBABILITY maps [N,C,H,W].
    Returns (boxes [N,8192,4], scores [N,8192], labels [N,8192] int32, counts [N] int32): device tensors, no host sync."""
    probs = [p.contiguous() for p in probs]
    regs = [r.contiguous() for r in box_regression]
    ctrs = [c.contiguous() for c in centerness]
    dev = probs[0].device
    n = geo.n_images
    hw = _upload_small(torch.tensor([[int(h), int(w)] for h, w in image_sizes], dtype=torch.int32), dev)
    boxes = torch.empty((n, PP_CAP, 4), device=dev, dtype=torch.float32)
    scores = torch.empty((n, PP_CAP), device=dev, dtype=torch.float32)
    labels = torch.empty((n, PP_CAP), device=dev, dtype=torch.int32)
    counts = torch.empty((n,), device=dev, dtype=torch.int32)
    ws_bytes = _lib.lib().scan_postprocess_workspace_bytes(n, len(probs))
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    call("scan_postprocess", geo.ref(), _ptr_array(probs), _ptr_array(regs), _ptr_array(ctrs), num_classes_fg, _ptr(hw),
         float(pre_nms_thresh), int(pre_nms_top_n), float(nms_thresh), int(post_top_n), float(min_size), _ptr(boxes), _ptr(scores),
         _ptr(labels), _ptr(counts), _ptr(ws), ws_bytes, _stream())
    return boxes, scores, labels, counts


# ----------------------------------------------------------------------------------------------------
# f3: CKA discriminator building blocks (csrc/cka.cu + the tower convolution kernels)
# ----------------------------------------------------------------------------------------------------
def thin_pack(geo, maps, c0, k, ld):
    """Channels [c0, c0 + k) of per-level NCHW maps -> rows [R, ld] (zero padded)."""
    maps = [m.contiguous() for m in maps]
    rows = torch.empty((geo.R, ld), device=maps[0].device, dtype=torch.float32)
    call("scan_thin_pack", geo.ref(), _ptr_array(maps), maps[0].shape[1], c0, k, _ptr(rows), ld, _stream())
    return rows


def thin_unpack(geo, rows, k_total, c0, k, scale=1.0):
    """rows [R, ld] columns [0, k) * scale -> per-level NCHW [N, k_total, H, W] tensors (other channels zero)."""
    outs = [torch.zeros((geo.n_images, k_total, h, w), device=rows.device, dtype=torch.float32) for h, w in geo.shapes]
    call("scan_thin_unpack", geo.ref(), _ptr(rows), rows.shape[1], k_total, c0, k, float(scale), _ptr_array(outs), _stream())
    return outs


def colsum(x, n_cols):
    """Column sums of the first n_cols columns of a contiguous [R, ld] matrix."""
    nbytes = _lib.lib().scan_colsum_workspace_bytes(x.shape[0], n_cols)
    ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
    out = torch.empty((n_cols,), device=x.device, dtype=torch.float32)
    call("scan_colsum", _ptr(x), x.shape[0], n_cols, x.shape[1], _ptr(out), _ptr(ws), nbytes, _stream())
    return out


class _GradScale(torch.autograd.Function):
    """layer.py:6-24 GradientReversalFunction: identity forward, dx = -lambda * grad (scan_scale)."""

    @staticmethod
    def forward(ctx, x, lambda_):
        ctx.lambda_ = float(lambda_)
        return x.view_as(x)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        g = g.contiguous() if g.is_contiguous() else nhwc_dense(g)
        out = torch.empty_like(g)
        if g.numel() % 4 == 0:
            call("scan_scale", _ptr(g), g.numel(), -ctx.lambda_, _ptr(out), _stream())
        else:     # odd-sized tails (tiny maps): pad through a flat copy
            flat = torch.zeros(((g.numel() + 3) // 4 * 4,), device=g.device, dtype=torch.float32)
            flat[:g.numel()] = g.reshape(-1)
            res = torch.empty_like(flat)
            call("scan_scale", _ptr(flat), flat.numel(), -ctx.lambda_, _ptr(res), _stream())
            out = res[:g.numel()].view(g.shape)
        return out, None


def grad_reverse(x, lambda_):
    return _GradScale.apply(x, lambda_)


def _pad256(n):
    return (n + 255) // 256 * 256


class _CkaClassMaps(torch.autograd.Function):
    """The per-class loop of FCOSDiscriminator_con.forward (fcos_head_discriminator_con.py:100-123) for ALL classes at once:
         h      = relu(conv3x3([x | maps]; w1) + b1)       w1 [C*128, 256 + C, 3, 3] block structured (class c: its 128 rows see
                                                            the 256 feature columns and map column c)
         logits = conv3x3(h; w2) + b2                       w2 [C, C*128, 3, 3] block diagonal
         loss   = class-weighted BCE-with-logits(logits, target; weights = maps)
    x_rows [R, 256], maps32 [R, 32] (maps of classes 1..C in columns 0..C-1).  Returns the scalar loss."""

    @staticmethod
    def forward(ctx, geo, x_rows, maps32, w1, b1, w2, b2, target, n_cls):
        precise = CONV["precise"]
        dev = x_rows.device
        hc = w1.shape[0]                       # C * 128
        hp = _pad256(hc)
        lo = (lambda t: tf32_residual(t)) if precise else (lambda t: None)
        w1_hi, w1_lo = conv3x3_pack(w1, False, precise)
        h = torch.zeros((geo.R, hp), device=dev, dtype=torch.float32) if hp != hc else torch.empty((geo.R, hp), device=dev, dtype=torch.float32)
        x_lo, m_lo = lo(x_rows), lo(maps32)
        call("scan_conv3x3_rows2", geo.ref(), _ptr(x_rows), _ptr(x_lo), x_rows.shape[1], _ptr(maps32), _ptr(m_lo), 32, _ptr(w1_hi),
             _ptr(w1_lo), hc, _ptr(b1), None, None, 1, _ptr(h), hp, CONV["cta_group"], _stream())
        # the second convolution reads h with its padded width: pad w2's input channels with zero columns
        w2p = w2 if hp == hc else torch.cat([w2, w2.new_zeros((w2.shape[0], hp - hc, 3, 3))], dim=1)
        w2_hi, w2_lo = conv3x3_pack(w2p, False, precise)
        logits = torch.empty((geo.R, 32), device=dev, dtype=torch.float32)
        h_lo = lo(h)
        call("scan_conv3x3_rows2", geo.ref(), _ptr(h), _ptr(h_lo), hp, None, None, 0, _ptr(w2_hi), _ptr(w2_lo), n_cls, _ptr(b2), None, None,
             0, _ptr(logits), 32, CONV["cta_group"], _stream())
        nbytes = _lib.lib().scan_cka_bce_workspace_bytes()
        ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
        loss = torch.empty((1,), device=dev, dtype=torch.float32)
        inv = torch.empty((16,), device=dev, dtype=torch.float32)
        call("scan_cka_bce_fwd", _ptr(logits), 32, _ptr(maps32), 32, geo.R, n_cls, float(target), _ptr(loss), _ptr(inv), _ptr(ws), nbytes,
             _stream())
        ctx.geo, ctx.precise, ctx.target, ctx.n_cls, ctx.hc, ctx.hp = geo, precise, float(target), n_cls, hc, hp
        ctx.save_for_backward(x_rows, maps32, w1, w2p, h, logits, inv)
        ctx.lo = (x_lo, m_lo, h_lo)
        return loss.reshape(())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_loss):
        geo, precise, n_cls, hc, hp = ctx.geo, ctx.precise, ctx.n_cls, ctx.hc, ctx.hp
        x_rows, maps32, w1, w2p, h, logits, inv = ctx.saved_tensors
        x_lo, m_lo, h_lo = ctx.lo
        dev = x_rows.device
        lo = (lambda t: tf32_residual(t)) if precise else (lambda t: None)
        d_loss = d_loss.reshape(1).contiguous().float()
        dl32 = torch.empty((geo.R, 32), device=dev, dtype=torch.float32)
        dl256 = torch.zeros((geo.R, 256), device=dev, dtype=torch.float32)
        call("scan_cka_bce_bwd", _ptr(logits), 32, _ptr(maps32), 32, geo.R, n_cls, ctx.target, _ptr(inv), _ptr(d_loss), _ptr(dl32),
             _ptr(dl256), 256, _stream())
        d_b2 = colsum(dl32, n_cls)
        # second convolution: weight gradient (dY padded to 256 columns for the 256-wide MMA tile) and data gradient through h's ReLU
        dl32_lo, dl256_lo = lo(dl32), lo(dl256)
        d_w2_full = torch.empty((256, hp, 3, 3), device=dev, dtype=torch.float32)
        conv3x3_wgrad_raw(geo, h, dl256, x_lo=h_lo, dy_lo=dl256_lo, out=d_w2_full)
        d_w2p = d_w2_full[:n_cls]
        w2t_hi, w2t_lo = conv3x3_pack(w2p, True, precise)      # rows = h channels (hp), cols = classes (padded to 32)
        d_pre = torch.zeros((geo.R, hp), device=dev, dtype=torch.float32) if hp != hc else torch.empty((geo.R, hp), device=dev, dtype=torch.float32)
        call("scan_conv3x3_rows2", geo.ref(), _ptr(dl32), _ptr(dl32_lo), 32, None, None, 0, _ptr(w2t_hi), _ptr(w2t_lo), hc, None, None,
             _ptr(h), 0, _ptr(d_pre), hp, CONV["cta_group"], _stream())
        d_b1 = colsum(d_pre, hc)
        d_pre_lo = lo(d_pre)
        # first convolution: gradients wrt the features, the class maps and the block-structured weight
        w1p = w1 if hp == hc else torch.cat([w1, w1.new_zeros((hp - hc,) + tuple(w1.shape[1:]))], dim=0)
        w1x_hi, w1x_lo = conv3x3_pack(w1p[:, :256], True, precise)
        d_x = conv3x3_rows_raw(geo, d_pre, w1x_hi, 256, x_lo=d_pre_lo, packed_lo=w1x_lo)
        w1m_hi, w1m_lo = conv3x3_pack(w1p[:, 256:], True, precise)
        d_maps32 = torch.zeros((geo.R, 32), device=dev, dtype=torch.float32)
        call("scan_conv3x3_rows2", geo.ref(), _ptr(d_pre), _ptr(d_pre_lo), hp, None, None, 0, _ptr(w1m_hi), _ptr(w1m_lo), n_cls, None, None,
             None, 0, _ptr(d_maps32), 32, CONV["cta_group"], _stream())
        d_w1 = torch.empty_like(w1.contiguous())
        d_w1x_full = torch.empty((hp, 256, 3, 3), device=dev, dtype=torch.float32)
        conv3x3_wgrad_raw(geo, x_rows, d_pre, x_lo=x_lo, dy_lo=d_pre_lo, out=d_w1x_full)
        maps256 = torch.zeros((geo.R, 256), device=dev, dtype=torch.float32)
        maps256[:, :32] = maps32
        d_w1m_full = torch.empty((hp, 256, 3, 3), device=dev, dtype=torch.float32)
        conv3x3_wgrad_raw(geo, maps256, d_pre, x_lo=lo(maps256), dy_lo=d_pre_lo, out=d_w1m_full)
        d_w1[:, :256] = d_w1x_full[:hc]
        d_w1[:, 256:] = d_w1m_full[:hc, :n_cls]
        d_w2 = d_w2p[:, :hc].contiguous()
        return None, d_x, d_maps32, d_w1, d_b1, d_w2, d_b2, None, None


def cka_class_maps_loss(geo, x_rows, maps32, w1, b1, w2, b2, target, n_cls):
    if x_rows.shape[1] != C or maps32.shape[1] != 32 or n_cls > 16:
        raise RuntimeError("cka_class_maps_loss: [R,256] features, [R,32] class maps, at most 16 classes")
    if not x_rows.is_cuda:
        raise RuntimeError("cka_class_maps_loss needs CUDA tensors (no CPU fallback)")
    return _CkaClassMaps.apply(geo, x_rows.contiguous(), maps32.contiguous(), w1.contiguous(), b1.contiguous(), w2.contiguous(),
                               b2.contiguous(), target, n_cls)


class _ThinPackMaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geo, maps, c0, k):
        ctx.geo, ctx.c0, ctx.k, ctx.k_total = geo, c0, k, maps.shape[1]
        return thin_pack(geo, [maps], c0, k, 32)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_rows):
        return None, thin_unpack(ctx.geo, d_rows.contiguous(), ctx.k_total, ctx.c0, ctx.k)[0], None, None


def thin_pack_maps(geo, maps, c0, k):
    """Single-level [N,K,H,W] class maps -> rows [R,32] holding channels [c0, c0 + k) (differentiable)."""
    if not maps.is_cuda or maps.dtype != torch.float32 or k > 32:
        raise RuntimeError("thin_pack_maps expects CUDA fp32 maps with at most 32 selected channels (no CPU fallback)")
    return _ThinPackMaps.apply(geo, maps, c0, k)


# ----------------------------------------------------------------------------------------------------
# head_out on the tower kernels: relu(conv3x3(cat([features, act_maps], 1)) + bias) without the concatenation (a12 / f1)
# ----------------------------------------------------------------------------------------------------
class _HeadOut(torch.autograd.Function):
    """y = relu(conv3x3([F | maps]; W) + b) in ONE scan_conv3x3_rows2 launch (two input tensors, bias + ReLU epilogue), or -- when
    the feature half u = conv3x3(F; W[:, :256]) has been enqueued earlier (multi-GPU: it hides the prototype all-reduce) --
    y = relu(conv3x3(maps; W[:, 256:]) + u + b) with u as the epilogue's addend.
    Backward: d_pre = dy * [y > 0] and d_b (scan_add_relu_bwd), d_maps by the N = 32 instantiation of the convolution kernel,
    d_F / d_W[:, :256] (fused form only) and d_W[:, 256:] by the data- and weight-gradient kernels."""

    @staticmethod
    def forward(ctx, geo, weight, bias, fused, n_levels, *tensors):
        # n_levels > 0: `first` = that many per-level tensors (fused: the feature levels; else: the early feature half u);
        # n_levels = 0 / -1: `first` = ONE [R,256] rows matrix (fused form); -1 additionally hands the rows on as an alias (last
        # output) whose gradient -- from consumers that run after head_out -- enters the data-gradient kernel as its addend
        precise = CONV["precise"]
        rows_in = n_levels <= 0
        ctx.rows_in, ctx.through = rows_in, n_levels < 0
        n_first = 1 if rows_in else n_levels
        first = tensors[:n_first]
        acts = [a.contiguous() for a in tensors[n_first:]]
        k = acts[0].shape[1]
        dev = acts[0].device
        lo = (lambda t: tf32_residual(t)) if precise else (lambda t: None)
        maps32 = torch.empty((geo.R, 32), device=dev, dtype=torch.float32)
        call("scan_thin_pack", geo.ref(), _ptr_array(acts), k, 0, k, _ptr(maps32), 32, _stream())
        m_lo = lo(maps32)
        y_rows = torch.empty((geo.R, C), device=dev, dtype=torch.float32)
        bias = bias.contiguous()
        first_rows = first[0].contiguous() if rows_in else _rows_of_levels(geo, list(first))
        f_lo = None
        if fused:
            hi, wlo = conv3x3_pack(weight, False, precise)
            f_lo = lo(first_rows)
            call("scan_conv3x3_rows2", geo.ref(), _ptr(first_rows), _ptr(f_lo), C, _ptr(maps32), _ptr(m_lo), 32, _ptr(hi), _ptr(wlo), C,
                 _ptr(bias), None, None, 1, _ptr(y_rows), C, CONV["cta_group"], _stream())
        else:
            hi, wlo = conv3x3_pack(weight[:, C:], False, precise)
            call("scan_conv3x3_rows2", geo.ref(), _ptr(maps32), _ptr(m_lo), 32, None, None, 0, _ptr(hi), _ptr(wlo), C, _ptr(bias),
                 _ptr(first_rows), None, 1, _ptr(y_rows), C, CONV["cta_group"], _stream())
        ctx.geo, ctx.precise, ctx.fused, ctx.n_levels, ctx.k = geo, precise, fused, n_levels, k
        ctx.save_for_backward(weight, y_rows, first_rows if fused else y_rows, maps32, *acts)
        ctx.f_lo = f_lo
        if ctx.through:
            ctx.set_materialize_grads(False)
            return tuple(level_views(geo, y_rows)) + (first[0],)
        return tuple(level_views(geo, y_rows))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *d_levels):
        geo, precise, fused, n_levels, k = ctx.geo, ctx.precise, ctx.fused, ctx.n_levels, ctx.k
        d_alias = None
        if ctx.through:
            d_alias, d_levels = d_levels[-1], d_levels[:-1]
            if all(g is None for g in d_levels):
                return (None,) * 5 + (d_alias,) + (None,) * len(ctx.saved_tensors[4:])
        weight, y_rows, f_rows, maps32 = ctx.saved_tensors[:4]
        acts = list(ctx.saved_tensors[4:])
        dev = y_rows.device
        lo = (lambda t: tf32_residual(t)) if precise else (lambda t: None)
        views = level_views(geo, y_rows)
        dys = [nhwc_dense(g) if g is not None else torch.zeros_like(v) for g, v in zip(d_levels, views)]
        d_pre = torch.empty_like(y_rows)
        d_bias = torch.empty((C,), device=dev, dtype=torch.float32)
        ws = torch.empty((_lib.lib().scan_gn_workspace_bytes(geo.ref()),), device=dev, dtype=torch.uint8)
        call("scan_add_relu_bwd", geo.ref(), _ptr_array(dys), _ptr(y_rows), _ptr(d_pre), _ptr(d_bias), _ptr(ws), ws.numel(), _stream())
        d_pre_lo = lo(d_pre)
        # d(maps): thin data gradient, back to per-level NCHW.  d_pre is read ONCE by a one-tap launch of the convolution kernel
        # against the [K * 9, 256] weight slice (what each pixel sends to its nine neighbours), then gathered over the taps; more than 14 classes take the
        # N = 32 instantiation of the convolution kernel (nine shifted reads of d_pre)
        d_maps32 = torch.empty((geo.R, 32), device=dev, dtype=torch.float32)
        if 9 * k <= 128:
            w_taps = torch.zeros((256, C), device=dev, dtype=torch.float32)
            w_taps[:9 * k] = weight[:, C:].permute(1, 2, 3, 0).reshape(9 * k, C)
            sent = torch.empty((geo.R, 128), device=dev, dtype=torch.float32)
            call("scan_conv1x1_rows", geo.ref(), _ptr(d_pre), _ptr(d_pre_lo), C, _ptr(w_taps), _ptr(lo(w_taps)), 9 * k, None, 0, _ptr(sent),
                 128, CONV["cta_group"], _stream())
            call("scan_thin_gather", geo.ref(), _ptr(sent), 128, k, _ptr(d_maps32), _stream())
        else:
            hi, wlo = conv3x3_pack(weight[:, C:], True, precise)
            call("scan_conv3x3_rows2", geo.ref(), _ptr(d_pre), _ptr(d_pre_lo), C, None, None, 0, _ptr(hi), _ptr(wlo), k, None, None, None, 0,
                 _ptr(d_maps32), 32, CONV["cta_group"], _stream())
        d_acts = [torch.empty((geo.n_images, k, h, w), device=dev, dtype=torch.float32) for h, w in geo.shapes]
        call("scan_thin_unpack", geo.ref(), _ptr(d_maps32), 32, k, 0, k, 1.0, _ptr_array(d_acts), _stream())
        # weight gradient of the map columns: tap-spread maps x d_pre as one MN-major GEMM (scan_thin_wgrad); more than 14 classes
        # fall back to the 256-wide kernel on a zero-padded copy of the maps
        d_w = torch.zeros_like(weight, memory_format=torch.contiguous_format)
        if 9 * k <= 128:
            nbytes = _lib.lib().scan_thin_wgrad_workspace_bytes(geo.ref(), int(precise))
            ws2 = _workspace("thin_wgrad", nbytes, dev)
            d_wm = d_w[:, C:]
            s = d_wm.stride()
            call("scan_thin_wgrad", geo.ref(), _ptr(maps32), _ptr(d_pre), _ptr(d_pre_lo), k, _ptr(d_wm), s[0], s[1], s[2], s[3], _ptr(ws2),
                 nbytes, _stream())
        else:
            maps256 = torch.empty((geo.R, C), device=dev, dtype=torch.float32)
            call("scan_thin_pack", geo.ref(), _ptr_array(acts), k, 0, k, _ptr(maps256), C, _stream())
            d_wm = torch.empty((C, C, 3, 3), device=dev, dtype=torch.float32)
            conv3x3_wgrad_raw(geo, maps256, d_pre, x_lo=lo(maps256), dy_lo=d_pre_lo, out=d_wm)
            d_w[:, C:] = d_wm[:, :k]
        if fused:
            conv3x3_wgrad_raw(geo, f_rows, d_pre, x_lo=ctx.f_lo if precise else None, dy_lo=d_pre_lo, out=d_w[:, :C])
            hi, wlo = conv3x3_pack(weight[:, :C], True, precise)
            if d_alias is not None and not d_alias.is_contiguous():
                d_alias = d_alias.contiguous()
            d_f = conv3x3_rows_raw(geo, d_pre, hi, C, addend=d_alias, x_lo=d_pre_lo, packed_lo=wlo)
            d_first = (d_f,) if ctx.rows_in else tuple(level_views(geo, d_f))
        else:
            d_first = tuple(level_views(geo, d_pre))        # the early feature half receives d_pre unchanged
        return (None, d_w, d_bias, None, None) + d_first + tuple(d_acts)


def head_out_levels(geo, weight, bias, acts, features=None, us=None, rows=None, through=False):
    """head_out's single convolution + ReLU on the tower kernels; exactly one of `features` (fused form, per-level tensors), `rows`
    (fused form, the [R,256] rows matrix; through=True also returns an alias of it as last element for later consumers) and `us`
    (early feature half)."""
    if (features is None) + (us is None) + (rows is None) != 2:
        raise RuntimeError("head_out_levels takes the feature levels, the rows matrix or the early feature half")
    if rows is not None:
        k = acts[0].shape[1]
        if weight.shape[0] != C or weight.shape[1] != C + k or k > 32 or rows.shape[1] != C:
            raise RuntimeError("head_out_levels is built for a [256, 256 + K <= 32, 3, 3] weight")
        if not rows.is_cuda:
            raise RuntimeError("head_out_levels needs CUDA fp32 tensors (no CPU fallback)")
        out = list(_HeadOut.apply(geo, weight, bias, True, -1 if through else 0, rows, *acts))
        return (out[:-1], out[-1]) if through else out
    k = acts[0].shape[1]
    if weight.shape[0] != C or weight.shape[1] != C + k or k > 32:
        raise RuntimeError("head_out_levels is built for a [256, 256 + K <= 32, 3, 3] weight")
    first = list(features if features is not None else us)
    for t in first + list(acts):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError("head_out_levels needs CUDA fp32 tensors (no CPU fallback)")
    return list(_HeadOut.apply(geo, weight, bias, features is not None, len(first), *first, *acts))
