"""B200-native condgraph middle head: host-side mirror of the reference's GRAPHModule.

Same constructor (`build_condgraph(cfg, in_channels)`), same forward signature and 4-tuple result,
same `state_dict()` keys/shapes, same MODEL.MIDDLE_HEAD.* keys as
/root/reference/fcos_core/modeling/rpn/fcos/condgraph.py (GRAPHModule :122-669, build_condgraph :672) and the node
sampling of .../fcos/loss.py (PrototypeComputation :239-520) -- but the hot path runs in hand-written sm_100a kernels
behind the C ABI of include/scan_b200.h:

    head_in (torch/cuDNN, out of the kernel scope: SURVEY §8a a1)
      -> scan_pack_rows            NCHW -> rows [R,256] (level-first, image-major: the reference's flattening)
      -> scan_fcos_assign          K1  GT -> location assignment                      (loss.py:262-343)
      -> scan_sample_nodes         K1  node index generation, balanced negatives      (loss.py:430-458 / 497-516)
      -> scan_gather_rows          K1  node rows
      -> scan_attn_fwd/bwd         K3  affinity + softmax + aggregation               (transformer.py:5-34)
      -> scan_class_sums + scan_proto_update   K3  per-class means + paradigm EMA     (condgraph.py:395-398, 558-617)
      -> manifestation (cuDNN RNN / cuBLAS; tiny)                                     (condgraph.py:313-336)
      -> scan_condconv_fwd/bwd     K4  conditional conv + softmax + focal loss, tcgen05 (condgraph.py:619-629, 338-370)
      -> scan_dbscan_level         K2  target-domain sampling                         (loss.py:397-423)
    head_out (torch/cuDNN)

There is no CPU path: features must be CUDA tensors and libscan_b200.so must be loadable.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .config import CfgNode  # noqa: F401  (re-exported for users without yacs)


class PROTOTYPECounter(object):
    """condgraph.py:46-65: stop=True yields 0,1,..,cycle,cycle,...; stop=False cycles 0..cycle-1."""

    def __init__(self, cycle=3, stop=False):
        self.cycle = cycle
        self.counter = -1
        self.stop = stop

    def __call__(self, *args, **kwargs):
        if self.stop:
            if self.counter != self.cycle:
                self.counter += 1
            return self.counter
        self.counter += 1
        if self.counter == self.cycle:
            self.counter = 0
        return self.counter


class GRAPHHead(nn.Module):
    """condgraph.py:68-119: [Conv3x3 (+GN/IN/BN) + ReLU] x num_convs, shared over FPN levels (stays torch/cuDNN)."""

    def __init__(self, cfg, in_channels, out_channel, mode="in"):
        super().__init__()
        mh = cfg.MODEL.MIDDLE_HEAD
        if mode == "in":
            num_convs = mh.NUM_CONVS_IN
        elif mode == "out":
            num_convs = mh.NUM_CONVS_OUT
        else:
            num_convs = cfg.MODEL.FCOS.NUM_CONVS
        tower = []
        for _ in range(num_convs):
            tower.append(nn.Conv2d(in_channels, out_channel, kernel_size=3, stride=1, padding=1))
            if mode == "in":
                if mh.IN_NORM == "GN":
                    tower.append(nn.GroupNorm(32, in_channels))
                elif mh.IN_NORM == "IN":
                    tower.append(nn.InstanceNorm2d(in_channels))
                elif mh.IN_NORM == "BN":
                    tower.append(nn.BatchNorm2d(in_channels))
            tower.append(nn.ReLU())
        self.add_module("middle_tower", nn.Sequential(*tower))
        for layer in self.middle_tower.modules():
            if isinstance(layer, nn.Conv2d):
                nn.init.normal_(layer.weight, std=0.01)
                nn.init.constant_(layer.bias, 0)

    def forward(self, x):
        return [self.middle_tower(f) for f in x]

    def forward_levels(self, geo, levels):
        """The same tower on channels-last tensors: cuDNN runs its NHWC kernels without layout conversions and every
        GroupNorm + ReLU pair is the scan_gn_relu kernel over all levels at once.  Returns per-level
        [N,C,H,W] tensors; after a GroupNorm they are adjacent views of one rows buffer (ops.join_rows is then free)."""
        layers = list(self.middle_tower)
        h, i = list(levels), 0
        while i < len(layers):
            conv = layers[i]
            i += 1
            if i + 1 < len(layers) and isinstance(layers[i], nn.GroupNorm) and isinstance(layers[i + 1], nn.ReLU) \
                    and layers[i].num_groups == 32:
                gn = layers[i]
                i += 2
                # bias-free convolution: the GroupNorm kernel adds the bias and returns its gradient as a by-product
                # (and takes its statistics from the convolution's epilogue when that is the tcgen05 kernel)
                outs, stats = _tower_conv(geo, conv.weight, None, h, gn=(conv.bias, gn.eps))
                h = ops.gn_relu_levels(geo, gn.weight, gn.bias, gn.eps, outs, conv_bias=conv.bias, stats=stats)
            else:   # IN / BN variants and the norm-free head_out: torch modules (not used by the shipped configs' head_in)
                w = conv.weight.contiguous(memory_format=torch.channels_last)
                outs = [F.conv2d(ops.nhwc_dense(x), w, conv.bias, padding=1) for x in h]
                while i < len(layers) and not isinstance(layers[i], nn.Conv2d):
                    outs = [layers[i](o) for o in outs]
                    i += 1
                h = outs
        return h


TOWERS = {"impl": os.environ.get("SCAN_B200_TOWERS", "scan")}     # "scan" (csrc/tower.cu) | "cudnn"


def _tower_conv(geo, weight, weight_cl, levels, gn=None):
    """Bias-free 3x3 tower convolution of all levels: the tcgen05 implicit GEMM of csrc/tower.cu (f1) for the 256 -> 256 layers of
    every shipped config; other widths, and SCAN_B200_TOWERS=cudnn (the A/B switch of tools/ and bench.py), go to cuDNN's
    channels-last kernels.
    gn = (conv_bias, eps) of a following GroupNorm(32): returns (levels, stats | None), the statistics being a by-product of the
    tcgen05 kernel's epilogue (ops.CONV["gn_stats"]); without gn: the levels."""
    own = TOWERS["impl"] == "scan" and weight.shape[0] % 256 == 0 and weight.shape[1] % 256 == 0
    if own and gn is not None and ops.CONV["gn_stats"] and weight.shape[0] == 256:
        return ops.conv3x3_levels(geo, weight, list(levels), gn=gn)
    if own:
        outs = ops.conv3x3_levels(geo, weight, list(levels))
    else:
        if weight_cl is None:
            weight_cl = weight.contiguous(memory_format=torch.channels_last)
        outs = [F.conv2d(ops.nhwc_dense(x), weight_cl, None, padding=1) for x in levels]
    return outs if gn is None else (outs, None)


class MultiHeadAttention(nn.Module):
    """Parameters of layers/transformer.py:36-90; the arithmetic runs in scan_qkv_fwd / scan_attn_fwd / scan_attn_out_ln_fwd
    and their backward entry points (no torch / cuBLAS op on this path)."""

    def __init__(self, model_dim=256, num_heads=4, dropout=0.1):
        super().__init__()
        self.dim_per_head = model_dim // num_heads
        self.num_heads = num_heads
        self.linear_k = nn.Linear(model_dim, model_dim)
        self.linear_v = nn.Linear(model_dim, model_dim)
        self.linear_q = nn.Linear(model_dim, model_dim)
        self.linear_final = nn.Linear(model_dim, model_dim)
        self.layer_norm = nn.LayerNorm(model_dim)
        self.p_drop = dropout

    def forward(self, x):
        """x [M,256] (the reference passes (key, value, query) = (x, x, x), condgraph.py:392)."""
        p = self.p_drop if self.training else 0.0
        seed = 0
        if p > 0:
            # one 63-bit draw per call from torch's default CPU generator: the masks follow torch.manual_seed / get_rng_state /
            # fork_rng and a checkpoint's RNG state like nn.Dropout does, and differ per rank when the ranks' generators do
            seed = int(torch.randint(0, 0x7FFFFFFFFFFFFFFF, (1,), dtype=torch.int64).item())
        # projections, attention core, linear_final + dropout + residual LayerNorm: four tcgen05 launches, one autograd node
        return ops.graph_attention(x, self, p, seed)


def sim_matrix(a, b, eps=1e-8):
    """condgraph.py:35-43."""
    a_n, b_n = a.norm(dim=1)[:, None], b.norm(dim=1)[:, None]
    return torch.mm(a / torch.clamp(a_n, min=eps), (b / torch.clamp(b_n, min=eps)).t())


class GRAPHModule(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.cfg = cfg.clone()
        mh = cfg.MODEL.MIDDLE_HEAD
        self.debug_cfg = cfg.MODEL.DEBUG_CFG
        if self.debug_cfg:
            raise RuntimeError("MODEL.DEBUG_CFG (t-SNE / map dumps with os._exit) is out of scope")
        self.with_bg_proto = bool(mh.PROTO_WITH_BG)
        self.with_bias_dc = bool(mh.COND_WITH_BIAS)
        self.with_concated_maps = bool(mh.CAT_ACT_MAP)
        self.with_shortcut_GCNs = bool(mh.GCN_SHORTCUT)
        self.with_global_gcn = bool(mh.GLOBAL_GCN)
        self.with_self_training = bool(mh.GCN_SELF_TRAINING)
        self.fpn_strides = list(cfg.MODEL.FCOS.FPN_STRIDES)
        self.num_classes_fg = cfg.MODEL.FCOS.NUM_CLASSES - 1
        self.used_num_classes = self.num_classes_fg + int(self.with_bg_proto)
        self.transfer_cfg = tuple(mh.TRANSFER_CFG)
        self.act_loss_cfg = mh.ACT_LOSS
        self.GCN_norm_cfg = mh.GCN_EDGE_NORM
        self.GCN_out_act_cfg = mh.GCN_OUT_ACTIVATION
        self.lamda1, self.lamda2 = mh.GCN_LOSS_WEIGHT, mh.ACT_LOSS_WEIGHT
        self.lamda3, self.lamda4 = mh.CON_LOSS_WEIGHT, mh.GCN_LOSS_WEIGHT_TG
        self.use_rnn = mh.USE_RNN
        self.prototype_iter = mh.PROTO_ITER
        self.cosine_update = bool(mh.COSINE_UPDATE_ON)
        self.target_sampling = mh.TARGET_SAMPLING_CFG
        self.dbscan_eps = float(mh.DBSCAN_EPS)
        self.dbscan_thr = float(mh.DBSCAN_THR)
        self.plabel_th = float(cfg.SOLVER.MIDDLE_HEAD.PLABEL_TH[0])
        # DBSCAN workspace capacity in points per level (adjacency = cap^2 / 8 bytes: 0.5 GB at 65 536, 1.25 GB at 100 000).
        # Not a limit: a level that selects more points grows the workspace and is re-run (see _sample_target); the
        # optional key MODEL.MIDDLE_HEAD.DBSCAN_MAX_POINTS only presets the initial size.
        self.dbscan_cap = int(getattr(mh, "DBSCAN_MAX_POINTS", 65536))
        self.dbscan_cap_limit = 700000   # 61 GB of adjacency bits: beyond this the O(n^2) clustering itself is hopeless
        self.dbscan_streams = True       # one side stream per FPN level for the DBSCAN launch sequences (see _dbscan_levels)
        self._side_streams = []
        self.record = False              # keep intermediate results of the last call in self.last (parity tests, bench stats)
        channel = mh.PROTO_CHANNEL
        hidden = mh.COND_HIDDEN_CHANNEL
        if channel != ops.C or in_channels != ops.C:
            raise RuntimeError("the sm_100a kernels are built for 256 channels")
        if self.used_num_classes > 16:
            raise RuntimeError("used_num_classes > 16 is not supported by the conditional-conv kernel")
        if self.act_loss_cfg == "sigmoidFL" and self.used_num_classes != 2:
            raise RuntimeError("sigmoidFL is hard-wired to 2 classes in the reference (condgraph.py:362-363)")

        self.head_in = GRAPHHead(cfg, in_channels, in_channels, mode="in")
        if self.prototype_iter == 1:
            self.register_buffer("prototype", torch.randn(self.used_num_classes, channel))
        else:
            self.register_buffer("prototype", torch.randn(self.used_num_classes, channel, self.prototype_iter))
        if self.with_concated_maps:
            self.head_out = GRAPHHead(cfg, in_channels + self.used_num_classes, in_channels, mode="out")
        self.proto_cls_hidden = nn.Linear(mh.GCN2_OUT_CHANNEL, 512)
        self.proto_cls = nn.Linear(512, self.used_num_classes)
        if self.with_global_gcn:
            self.multihead_attn = MultiHeadAttention(256, 4, dropout=0.1)
        else:
            if self.GCN_norm_cfg not in ("NO", "cosine_detached"):
                # 'softmax' / 'cosine' reference the never-defined edge_project_u/v (condgraph.py:289, 296)
                raise AttributeError("GCN_EDGE_NORM %r needs edge_project_u/v, which the reference never defines"
                                     % (self.GCN_norm_cfg,))
            self.gcn_layer1 = nn.Linear(256, mh.GCN1_OUT_CHANNEL)
            self.gcn_layer2 = nn.Linear(mh.GCN1_OUT_CHANNEL, mh.GCN2_OUT_CHANNEL)
            for layer in (self.gcn_layer1, self.gcn_layer2):
                nn.init.normal_(layer.weight, std=0.01)
                nn.init.constant_(layer.bias, 0)
        if self.use_rnn:
            self.cond_nx1 = nn.Conv2d(512, 256, kernel_size=(self.prototype_iter, 1))
            self.cond_rnn = nn.RNN(256, 512, 2, nonlinearity="tanh")
            self.counter_rnn = PROTOTYPECounter(self.prototype_iter, stop=True)
        elif self.prototype_iter > 1:
            self.counter = PROTOTYPECounter(self.prototype_iter)
            self.cond_nx1 = nn.Conv2d(channel, hidden, kernel_size=(self.prototype_iter, 1))
            nn.init.normal_(self.cond_nx1.weight)
            nn.init.constant_(self.cond_nx1.bias, 0)
            self.cond_nx1_norm = nn.GroupNorm(32, hidden)
        else:
            self.cond_1 = nn.Linear(channel, hidden)
            nn.init.normal_(self.cond_1.weight, std=0.01)
            nn.init.constant_(self.cond_1.bias, 0)
        self.cond_2 = nn.Linear(hidden, 256 + int(self.with_bias_dc))
        for layer in (self.cond_2, self.proto_cls, self.proto_cls_hidden):
            nn.init.normal_(layer.weight, std=0.01)
            nn.init.constant_(layer.bias, 0)
        self.last = {}     # intermediate results of the last call, only filled when self.record is set
        self.dist_group = None   # set by scan_b200.dist.attach() to all-reduce the prototype sums

    # ------------------------------------------------------------------ manifestation (condgraph.py:313-336)
    def get_conded_weight(self):
        if self.use_rnn:
            # h = cond_rnn(prototype.permute(2, 0, 1)); kernel = einsum("pkc,ocp->ko", h, cond_nx1.weight[..., 0]) + bias
            return ops.manifest_rnn(self.prototype, self.cond_rnn, self.cond_nx1)
        if self.prototype_iter > 1:
            # cond_nx1 = Conv2d(256, hid, (P,1)) over [K,256,P,1] == Linear(256 P, hid) on the flattened paradigm rows; GroupNorm32
            # over each row + ReLU; cond_2 (condgraph.py:322-328): three tiny-batch kernels (csrc/rowsmlp.cu)
            k = self.prototype.shape[0]
            hcat = ops.rows_linear(self.prototype.reshape(k, -1), self.cond_nx1.weight, self.cond_nx1.bias)
            hcat = ops.rows_gn_relu(hcat, self.cond_nx1_norm.weight, self.cond_nx1_norm.bias, 32, self.cond_nx1_norm.eps)
            return ops.rows_linear(hcat, self.cond_2.weight, self.cond_2.bias)
        # PROTO_ITER == 1 (condgraph.py:330-334)
        return ops.rows_linear(ops.rows_linear(self.prototype, self.cond_1.weight, self.cond_1.bias, relu=True),
                               self.cond_2.weight, self.cond_2.bias)

    def _split_kernel(self, kernel_par):
        if self.with_bias_dc:
            return kernel_par[:, :-1].contiguous(), kernel_par[:, -1].contiguous()
        return kernel_par, None

    def _act_mode(self):
        return 0 if self.act_loss_cfg == "softmaxFL" else 1   # condgraph.py:344-346: anything else is sigmoid

    # ------------------------------------------------------------------ graph aggregation (condgraph.py:386-421)
    def _forward_gcns(self, pos_points, pos_labels):
        k = self.used_num_classes
        shift = 0 if self.with_bg_proto else 1
        if self.with_global_gcn:
            nodes = self.multihead_attn(pos_points)
            if self.with_shortcut_GCNs:
                nodes = nodes + pos_points
        else:
            nodes = self._local_gcn(pos_points, pos_labels, shift)
        # per-class means (differentiable: the target branch feeds them to the transfer losses) + the sum|count buffer
        means, packed = ops.class_means(nodes, pos_labels, k, shift)
        node_loss = ops.node_classifier_loss(nodes, self.proto_cls_hidden, self.proto_cls, pos_labels, shift, self.lamda1)
        return node_loss, packed, nodes, means

    def _local_gcn(self, pos_points, pos_labels, shift):
        """Per-class GCN (condgraph.py:262-302, 404-414): Adj = softmax(affinity).detach(); two graph convolutions per class,
        all on the scan_b200 GEMM / softmax / activation kernels (ops.local_gcn)."""
        return ops.local_gcn(pos_points, pos_labels, self.gcn_layer1, self.gcn_layer2, self.used_num_classes, shift,
                             self.GCN_norm_cfg, self.GCN_out_act_cfg, self.with_shortcut_GCNs)

    # ------------------------------------------------------------------ paradigm update (condgraph.py:304-311, 558-617)
    @torch.no_grad()
    def update_prototype_ensemble(self, packed, between=None):
        """between: optional callable enqueued after the all-reduce has been STARTED and before its result is awaited (the
        source pass puts the feature half of head_out there): the ranks are not in lockstep, so the collective is a
        rendezvous, and independent work queued behind its start hides the wait for the slower rank."""
        if self.dist_group is not None:
            import torch.distributed as dist
            work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.dist_group, async_op=True)
            if between is not None:
                between()
            work.wait()        # stream-level wait on the collective's stream (no host block for NCCL)
        elif between is not None:
            between()
        shift = False
        if self.use_rnn:
            it = self.counter_rnn()
            if it == self.prototype_iter:
                slot, shift = it - 1, True
            else:
                slot = it
        elif self.prototype_iter > 1:
            slot = self.counter()
        else:
            slot = 0
        return ops.proto_update(packed, self.prototype, slot, shift, self.cosine_update, 0.95)

    # ------------------------------------------------------------------ head_out (condgraph.py:379-384)
    def head_out_feature_half(self, features):
        """The 256 feature columns of head_out's first convolution (no dependence on the activation maps): can be enqueued early."""
        conv = list(self.head_out.middle_tower)[0]
        wf = conv.weight[:, :ops.C]
        geo = ops.Geometry.of(features, self.fpn_strides)
        return _tower_conv(geo, wf, None, features)

    def _head_out_on_tower_kernels(self):
        if not self.with_concated_maps:
            return False
        layers = list(self.head_out.middle_tower)
        conv = layers[0]
        return (TOWERS["impl"] == "scan" and len(layers) == 2 and isinstance(layers[1], nn.ReLU) and conv.weight.shape[0] == ops.C
                and conv.bias is not None)

    def features_post_processing(self, features, act_maps, us=None, rows=None, through=False):
        """head_out(cat([features, act_maps], 1)) without materialising the concatenation: the first
        convolution is split into its 256 feature columns (channels-last, NHWC kernels) and its K map columns.
        us: the feature half when the caller has already enqueued it (head_out_feature_half).
        rows: the [R,256] rows matrix of `features` as the autograd input of the fused form (the alias chain gather -> conditional
        conv -> head_out, whose gradients meet inside kernel epilogues instead of autograd's add kernels); through=True returns
        (levels, rows alias) for a consumer that comes after head_out."""
        if not self.with_concated_maps:
            return (features, rows) if through else features
        layers = list(self.head_out.middle_tower)
        conv = layers[0]
        if self._head_out_on_tower_kernels():
            # the shipped form, entirely on the tower kernels: one launch over [features | maps] with bias + ReLU in its epilogue
            geo = ops.Geometry.of(features, self.fpn_strides)
            if us is None and rows is not None:
                return ops.head_out_levels(geo, conv.weight, conv.bias, list(act_maps), rows=rows, through=through)
            if us is None:
                out = ops.head_out_levels(geo, conv.weight, conv.bias, list(act_maps), features=list(features))
            else:
                out = ops.head_out_levels(geo, conv.weight, conv.bias, list(act_maps), us=list(us))
            return (out, rows) if through else out
        if through:
            return self.features_post_processing(features, act_maps, us=us), rows
        wa = conv.weight[:, ops.C:].contiguous(memory_format=torch.channels_last)
        if us is None:
            us = self.head_out_feature_half(features)
        vs = [F.conv2d(a.contiguous(memory_format=torch.channels_last), wa, None, padding=1) for a in act_maps]
        if len(layers) == 2 and isinstance(layers[1], nn.ReLU):
            geo = ops.Geometry.of(features, self.fpn_strides)
            return ops.add_relu_levels(geo, conv.bias, us, vs)      # relu(u + v + bias), all levels in one launch
        outs = []
        for u, v in zip(us, vs):
            y = u + v + conv.bias.view(1, -1, 1, 1)
            for layer in layers[1:]:
                y = layer(y)
            outs.append(y)
        return outs

    # ------------------------------------------------------------------ branches
    def _forward_train_source(self, images, features, targets=None, return_maps=False, pre=None):
        geo = ops.Geometry.of(features, self.fpn_strides)
        dev = features[0].device
        rows = ops.join_rows(geo, features)
        if pre is None:
            boxes, box_labels, box_count, g_max = ops.pad_targets(targets, dev)
            labels = ops.fcos_assign(geo, boxes, box_labels, box_count, g_max)
            pre = (labels, ops.sample_nodes(geo, 0, self.with_bg_proto, labels=labels))
        labels, smp = pre
        self._record_nodes(geo, smp, labels=labels)
        # `rows` continues as an alias handed through the gather: the conditional conv's d_rows then meets d_nodes inside ONE
        # backward node (no zero-filled scatter target, no full-size gradient sum)
        pos_points, rows = ops.gather_rows_through(rows, smp.node_rows)
        node_loss, packed, _, _ = self._forward_gcns(pos_points, smp.node_labels)
        # the feature half of head_out does not depend on the paradigm: it is enqueued between the start of the prototype
        # all-reduce and the wait for its result (multi-GPU: hides the rendezvous with the slower rank)
        held = {}
        grad_on = torch.is_grad_enabled()

        def early():
            # single GPU: nothing to hide, head_out then runs as ONE fused launch after the conditional convolution
            if self.with_concated_maps and self._dist_world() > 1:
                with torch.set_grad_enabled(grad_on):      # update_prototype_ensemble itself runs under no_grad
                    held["us"] = self.head_out_feature_half(features)

        proto_batch = self.update_prototype_ensemble(packed, between=early)
        weight, bias = self._split_kernel(self.get_conded_weight())
        if self.record:
            self.last["prototype_batch"], self.last["conded_weight"] = proto_batch, weight
        with_loss = self.act_loss_cfg in ("softmaxFL", "sigmoidFL")
        # gather -> conditional conv -> head_out read the rows through a chain of aliases: each backward adds its gradient into the
        # buffer the later consumer wrote (scatter-add / accumulate_rows / the convolution's addend), no full-size autograd sums
        chain = self._head_out_on_tower_kernels() and "us" not in held and torch.is_grad_enabled()
        res = ops.condconv(geo, rows, weight, bias, self.used_num_classes, self._act_mode(), labels if with_loss else None, self.lamda2,
                           through=chain)
        acts, act_loss, flags = res[:3]
        if self.record:
            self.last["act_loss_flags"] = flags
        out = self.features_post_processing(features, acts, us=held.get("us"), rows=res[3] if chain else None)
        return out, (node_loss, 0), act_loss, acts

    def get_transfer_loss(self, tg_prototype, tg_nodes, tg_labels):
        """condgraph.py:457-498: NODES / PROTOTYPE KL and ADJ / ADJ_COMPLETE cosine losses in the
        scan_transfer_* kernels; the class-presence mask stays on the device (the reference's boolean indexing synchronises)."""
        return ops.transfer_loss(self.transfer_cfg, tg_nodes, tg_labels, tg_prototype, self.prototype.detach())

    def _sample_target(self, geo, rows, acts, between=None):
        """between: optional callable enqueued on the caller's stream after the per-level DBSCAN sequences have been forked
        and before they are joined (the target pass puts head_out there: it does not depend on the sampling)."""
        dev = rows.device
        k = self.used_num_classes
        pos_mask = torch.empty((geo.R,), device=dev, dtype=torch.uint8)
        plabel = torch.empty((geo.R,), device=dev, dtype=torch.int64)
        infos = []
        if self.target_sampling == "dbscan":
            infos = self._dbscan_levels(geo, rows, acts, pos_mask, plabel, between=between)
            between = None
        elif self.target_sampling == "score_threshold":
            # loss.py:479-481 (alternative sampler; torch ops, not a north-star kernel)
            for l, act in enumerate(acts):
                a, b = geo.row_off[l], geo.row_off[l + 1]
                flat = act.detach().permute(0, 2, 3, 1).reshape(-1, k)
                pos_mask[a:b] = (flat[:, 1:] > self.plabel_th).sum(dim=-1).bool().to(torch.uint8)
                plabel[a:b] = flat[:, 1:].argmax(dim=-1) + 1
        else:
            raise KeyError("unknown target labels!")   # 'mean_shift' / 'kmeans' samplers are out of scope
        if between is not None:
            between()
        smp = ops.sample_nodes(geo, 1, True, pos_mask=pos_mask, plabel=plabel)
        if infos:
            # ONE deferred status read for all levels (the sampling above already synchronised).  A level that selected more
            # points than the workspace holds wrote nothing: grow the workspace to what it asked for and redo the pass.
            info = torch.stack(infos).cpu()
            if bool((info[:, 4] != 0).any()):
                need = int(info[:, 0].max())
                if need > self.dbscan_cap_limit:
                    raise RuntimeError("DBSCAN: %d points at one level (limit %d: the adjacency alone would take %.0f GB)"
                                       % (need, self.dbscan_cap_limit, need * need / 8e9))
                self.dbscan_cap = min(self.dbscan_cap_limit, (need * 5 // 4 + 1023) // 1024 * 1024)
                infos = self._dbscan_levels(geo, rows, acts, pos_mask, plabel)
                smp = ops.sample_nodes(geo, 1, True, pos_mask=pos_mask, plabel=plabel)
                info = torch.stack(infos).cpu()
                if bool((info[:, 4] != 0).any()):
                    raise RuntimeError("DBSCAN workspace growth failed")
            if self.record:
                self.last["dbscan_info"] = info
                self.last["dbscan_masks"] = [m.bool() for m in geo.split_rows(pos_mask)]
        return smp

    def _dbscan_levels(self, geo, rows, acts, pos_mask, plabel, between=None):
        """scan_dbscan_level per FPN level into the shared pos_mask / plabel vectors; returns the per-level info records.
        The levels are independent (loss.py:464-518 loops over them), so each one is enqueued on its own side stream: the
        ~20 small launches of the coarse levels (n = 100 .. 3 000 points: pure launch latency) overlap the P3 level's
        tensor-core distance kernel instead of queueing behind it.  Fork / join with events; no host synchronisation.
        `between()` is called between fork and join: with it every level (P3 too) goes to a side stream and the caller's
        stream carries head_out meanwhile -- the sampling's host read then finds the GPU busy instead of draining it."""
        k = self.used_num_classes
        caps = [min(geo.n_images * (k - 1) * h * w, self.dbscan_cap) for h, w in geo.shapes]
        n_levels = len(geo.shapes)
        dev = rows.device
        rows_d = rows.detach()
        main = torch.cuda.current_stream(dev)
        first_side = 0 if (between is not None and self.dbscan_streams) else 1     # level index from which side streams are used
        if self.dbscan_streams:
            if len(self._side_streams) < n_levels:
                self._side_streams = [torch.cuda.Stream(device=dev) for _ in range(n_levels)]
            fork = torch.cuda.Event()
            fork.record(main)
        infos = [None] * n_levels
        joins = []
        # bench.py's per-entry timing: the levels overlap and the P3 level (the longest chain) bounds the fork -> join span:
        # its own stream carries the two timing events
        timing = ops.TIMING["on"]
        if timing:
            ops.TIMING["on"] = False
        for l in range(n_levels):
            a, b = geo.row_off[l], geo.row_off[l + 1]
            side = self._side_streams[l] if (self.dbscan_streams and l >= first_side) else None
            if side is not None:
                side.wait_event(fork)
            with torch.cuda.stream(side if side is not None else main):
                if timing and l == 0:
                    span0 = torch.cuda.Event(enable_timing=True)
                    span0.record()
                ws = ops.dbscan_workspace(caps[l], dev)      # allocated on (and recycled by) the stream that uses it
                _, infos[l] = ops.dbscan_level(rows_d[a:b], acts[l].detach(), self.dbscan_thr, self.dbscan_eps, caps[l],
                                               pos_mask[a:b], plabel[a:b], ws)
                if timing and l == 0:
                    span1 = torch.cuda.Event(enable_timing=True)
                    span1.record()
                    ops.TIMERS.append(("scan_dbscan_levels_span", span0, span1))
                if side is not None:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    joins.append(ev)
        if timing:
            ops.TIMING["on"] = True
        if between is not None:
            between()
        for ev in joins:
            main.wait_event(ev)
        return infos

    def _forward_train_target(self, images, features, targets=None, return_maps=False):
        geo = ops.Geometry.of(features, self.fpn_strides)
        rows = ops.join_rows(geo, features)
        weight, bias = self._split_kernel(self.get_conded_weight())   # identical at every level (condgraph.py:507)
        # alias chain conditional conv -> head_out -> node gather (see _forward_train_source)
        chain = self._head_out_on_tower_kernels() and torch.is_grad_enabled()
        res = ops.condconv(geo, rows, weight, bias, self.used_num_classes, self._act_mode(), through=chain)
        acts = res[0]
        # head_out needs the maps only: enqueue it between the fork and the join of the DBSCAN streams, ahead of the host read
        held = {}
        grad_on = torch.is_grad_enabled()

        def head_out():
            with torch.set_grad_enabled(grad_on):
                if chain:
                    held["out"], held["rows"] = self.features_post_processing(features, acts, rows=res[3], through=True)
                else:
                    held["out"] = self.features_post_processing(features, acts)

        smp = self._sample_target(geo, rows, acts, between=head_out)
        self._record_nodes(geo, smp)
        out = held["out"]
        if smp.n_nodes > 0 and (self.transfer_cfg[0] is not None or self.with_self_training):
            pos_points = ops.gather_rows(held.get("rows", rows), smp.node_rows)
            # the class means feed the transfer losses WITH gradient in the reference (condgraph.py:398, 526)
            node_loss, packed, nodes, tg_proto = self._forward_gcns(pos_points, smp.node_labels)
            node_loss = self.lamda4 * node_loss
            # with GLOBAL_GCN=False the reference has overwritten the sampled rows in place (condgraph.py:413)
            tl_nodes = pos_points if self.with_global_gcn else nodes
            transfer_loss = self.get_transfer_loss(tg_proto, tl_nodes, smp.node_labels)
            if transfer_loss is not None:
                transfer_loss = self.lamda3 * transfer_loss
            if self.with_self_training:
                return out, (node_loss, transfer_loss), None, acts
            return out, (None, transfer_loss), None, acts
        return out, None, None, acts

    def _forward_inference(self, images, features, targets=None, return_maps=False):
        geo = ops.Geometry.of(features, self.fpn_strides)
        rows = ops.join_rows(geo, features)
        weight, bias = self._split_kernel(self.get_conded_weight())
        acts, _, _ = ops.condconv(geo, rows, weight, bias, self.used_num_classes, self._act_mode())
        return self.features_post_processing(features, acts), None, None, acts

    def forward(self, images, features, targets=None, return_maps=False, mode="source", forward_target=False):
        features = list(features)
        if not features[0].is_cuda:
            raise RuntimeError("scan_b200.GRAPHModule runs on CUDA only (no CPU fallback)")
        geo = ops.Geometry.of(features, self.fpn_strides)
        source = self.training and targets and mode == "source"
        pre = None
        if source:
            # The FCOS assignment and the source node sampling depend on the targets and the level geometry only, not on the
            # features: do them FIRST.  Their host read (the node count M sizes the graph tensors) then happens while the GPU
            # is otherwise idle for ~50 us, and the rest of the source pass -- towers, graph aggregation, conditional conv --
            # is enqueued without another synchronisation (before: the read sat after head_in and drained the queue).
            # They run on a side stream: the read then waits for these three small launches only, NOT for whatever the main
            # stream still holds (the previous step's backward), so the host keeps enqueueing ahead of the GPU across steps.
            dev = features[0].device
            main, side = torch.cuda.current_stream(dev), self._pre_stream(dev)
            if any(t.bbox.is_cuda for t in targets):
                side.wait_stream(main)          # device-resident targets may have been produced on the main stream
            with torch.cuda.stream(side):
                boxes, box_labels, box_count, g_max = ops.pad_targets(targets, dev)
                labels = ops.fcos_assign(geo, boxes, box_labels, box_count, g_max)
                smp = ops.sample_nodes(geo, 0, self.with_bg_proto, labels=labels)
            main.wait_stream(side)
            for t in (boxes, box_labels, box_count, labels, smp.node_rows, smp.node_labels):
                t.record_stream(main)           # allocated on the side stream, consumed on the main one
            pre = (labels, smp)
        features = self.head_in.forward_levels(geo, ops.pack_levels(geo, features))
        self.last = {"features_in": features} if self.record else {}
        if source:
            return self._forward_train_source(images, features, targets, return_maps, pre=pre)
        elif self.training and mode == "target" and forward_target:
            return self._forward_train_target(images, features, targets=None, return_maps=return_maps)
        return self._forward_inference(images, features, targets=None, return_maps=return_maps)

    def _dist_world(self):
        if self.dist_group is None:
            return 1
        import torch.distributed as dist
        return dist.get_world_size(self.dist_group)

    def _pre_stream(self, dev):
        streams = self.__dict__.setdefault("_pre_streams", {})
        key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
        if key not in streams:
            streams[key] = torch.cuda.Stream(device=dev)
        return streams[key]

    # ------------------------------------------------------------------ bookkeeping for tests
    def _record_nodes(self, geo, smp, labels=None):
        if not self.record:
            return
        if labels is not None:
            self.last["labels"] = geo.split_rows(labels)
        if smp.n_nodes == 0:
            self.last["node_rows"] = None
            return
        g = smp.node_rows.long()
        off = torch.tensor(geo.row_off[:-1], device=g.device)
        lv = torch.bucketize(g, off, right=True) - 1
        self.last["node_level"] = lv
        self.last["node_rows"] = g - off[lv]
        self.last["node_labels"] = smp.node_labels
        self.last["sample_meta"] = smp.meta


def build_condgraph(cfg, in_channels):
    """Drop-in for fcos_core.modeling.rpn.fcos.condgraph.build_condgraph (condgraph.py:672-673)."""
    return GRAPHModule(cfg, in_channels)
