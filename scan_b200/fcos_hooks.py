"""The two places where the FCOS head touches the hot path.

* `apply_test_mode` -- TEST.MODE 'common' / 'light' / 'precision' ensembling of the classification maps with the
  activation maps (fcos_core/modeling/rpn/fcos/fcos.py:159-169; the post-processor's own sigmoid of
  inference.py:68 applies in 'common' mode only).
* `FCOSLossComputation` -- drop-in for fcos_core.modeling.rpn.fcos.loss.FCOSLossComputation (loss.py:25-230; built by
  `make_fcos_loss_evaluator(cfg)`, loss.py:235-237): the GT assignment kernel of the middle head (it is the identical
  `[L, G]` matching) plus one fused loss pass over the head's maps.
* `SigmoidFocalLoss` -- drop-in for fcos_core.layers.SigmoidFocalLoss (layers/sigmoid_focal_loss.py:56-77) whose
  CUDA path was `_C.sigmoid_focalloss_forward/backward`.
"""
import torch
from torch import nn

from . import ops


def apply_test_mode(mode, box_cls, act_maps):
    """box_cls: list of per-level logits [N,K-1,H,W] (ignored / may be None for 'light');
    returns the per-level class-probability maps the FCOS post-processor thresholds."""
    if mode not in ("common", "light", "precision"):
        raise KeyError("unknown TEST.MODE %r" % (mode,))
    # one launch for all levels ('light' hands out views of the activation maps: no kernel at all)
    return ops.ensemble_levels(mode, None if mode == "light" else box_cls, list(act_maps))


class SigmoidFocalLoss(nn.Module):
    def __init__(self, gamma, alpha):
        super().__init__()
        self.gamma = gamma
        self.alpha = alpha

    def forward(self, logits, targets):
        if not logits.is_cuda:
            raise RuntimeError("scan_b200.SigmoidFocalLoss runs on CUDA only")
        return ops.sigmoid_focal_loss(logits, targets, self.gamma, self.alpha).sum()

    def __repr__(self):
        return "%s(gamma=%s, alpha=%s)" % (self.__class__.__name__, self.gamma, self.alpha)


class FCOSLossComputation(object):
    """loss.py:25-230.  __call__(locations, box_cls, box_regression, centerness, targets) -> (cls_loss, reg_loss, centerness_loss).
    `locations` is accepted for API compatibility; the kernels derive every location from its index (condgraph.py:631-655)."""

    def __init__(self, cfg):
        self.gamma = float(cfg.MODEL.FCOS.LOSS_GAMMA)
        self.alpha = float(cfg.MODEL.FCOS.LOSS_ALPHA)
        self.fpn_strides = list(cfg.MODEL.FCOS.FPN_STRIDES)

    def __call__(self, locations, box_cls, box_regression, centerness, targets):
        if not box_cls[0].is_cuda:
            raise RuntimeError("scan_b200.FCOSLossComputation runs on CUDA only")
        geo = ops.Geometry.of(box_cls, self.fpn_strides)
        boxes, box_labels, box_count, g_max = ops.pad_targets(targets, box_cls[0].device)
        labels, reg_targets = ops.fcos_assign_reg(geo, boxes, box_labels, box_count, g_max)
        losses = ops.fcos_loss(geo, labels, reg_targets, list(box_cls), list(box_regression), list(centerness), self.gamma, self.alpha)
        return losses[0], losses[1], losses[2]


def make_fcos_loss_evaluator(cfg):
    """loss.py:235-237."""
    return FCOSLossComputation(cfg)
