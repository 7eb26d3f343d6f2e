"""The two places where the FCOS head touches the hot path.

* `apply_test_mode` -- TEST.MODE 'common' / 'light' / 'precision' ensembling of the classification maps with the
  activation maps (fcos_core/modeling/rpn/fcos/fcos.py:159-169; the post-processor's own sigmoid of
  inference.py:68 applies in 'common' mode only).
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
