"""The two places where the FCOS head touches the hot path.

* `apply_test_mode` -- TEST.MODE 'common' / 'light' / 'precision' ensembling of the classification maps with the
  activation maps (fcos_core/modeling/rpn/fcos/fcos.py:159-169; the post-processor's own sigmoid of
  inference.py:68 applies in 'common' mode only).
* `FCOSLossComputation` -- drop-in for fcos_core.modeling.rpn.fcos.loss.FCOSLossComputation (loss.py:25-230; built by
  `make_fcos_loss_evaluator(cfg)`, loss.py:235-237): the GT assignment kernel of the middle head (it is the identical
  `[L, G]` matching) plus one fused loss pass over the head's maps.
* `FCOSPostProcessor` -- drop-in for fcos_core.modeling.rpn.fcos.inference.FCOSPostProcessor (inference.py:25-194; built by
  `make_fcos_postprocessor(cfg)`, :197-214): candidate selection, top-k, box decoding, per-class NMS and the
  detections-per-image cap in four launches for all levels and images.
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


class FCOSPostProcessor(nn.Module):
    """inference.py:25-194.  forward(locations, box_cls, box_regression, centerness, image_sizes) -> list[BoxList] with fields
    "labels" (int64, 1-based) and "scores".  For TEST.MODE 'common' box_cls are logits (the reference applies the sigmoid here,
    :68); for 'light' / 'precision' they are the probabilities apply_test_mode produced."""

    def __init__(self, pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n, min_size, num_classes, mode="common",
                 fpn_strides=(8, 16, 32, 64, 128)):
        super().__init__()
        self.pre_nms_thresh, self.pre_nms_top_n, self.nms_thresh = pre_nms_thresh, pre_nms_top_n, nms_thresh
        self.fpn_post_nms_top_n, self.min_size, self.num_classes, self.mode = fpn_post_nms_top_n, min_size, num_classes, mode
        self.fpn_strides = list(fpn_strides)

    def forward(self, locations, box_cls, box_regression, centerness, image_sizes):
        from .structures import BoxList
        if not box_cls[0].is_cuda:
            raise RuntimeError("scan_b200.FCOSPostProcessor runs on CUDA only")
        geo = ops.Geometry.of(box_cls, self.fpn_strides)
        probs = list(box_cls)
        if self.mode == "common":
            probs = ops.ensemble_levels_common(probs)
        boxes, scores, labels, counts = ops.postprocess(geo, probs, list(box_regression), list(centerness), image_sizes,
                                                        box_cls[0].shape[1], self.pre_nms_thresh, self.pre_nms_top_n, self.nms_thresh,
                                                        self.fpn_post_nms_top_n, self.min_size)
        counts = counts.cpu().tolist()           # the only host read: the result is a list of variable-length Python objects
        out = []
        for i, (h, w) in enumerate(image_sizes):
            k = counts[i]
            bl = BoxList(boxes[i, :k], (int(w), int(h)), mode="xyxy")
            bl.add_field("labels", labels[i, :k].long())
            bl.add_field("scores", scores[i, :k])
            out.append(bl)
        return out


def make_fcos_postprocessor(config):
    """inference.py:197-214."""
    return FCOSPostProcessor(pre_nms_thresh=config.MODEL.FCOS.INFERENCE_TH, pre_nms_top_n=config.MODEL.FCOS.PRE_NMS_TOP_N,
                             nms_thresh=config.MODEL.FCOS.NMS_TH, fpn_post_nms_top_n=config.TEST.DETECTIONS_PER_IMG, min_size=0,
                             num_classes=config.MODEL.FCOS.NUM_CLASSES, mode=config.TEST.MODE,
                             fpn_strides=config.MODEL.FCOS.FPN_STRIDES)
