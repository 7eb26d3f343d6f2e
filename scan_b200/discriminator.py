"""Drop-in for fcos_core.modeling.discriminator.FCOSDiscriminator_con, the CKA discriminator that consumes the middle head's
features AND its activation maps through a gradient reversal layer (fcos_head_discriminator_con.py:12-127, layer.py:6-45) --
SURVEY §8 row f3.  Same constructor, parameter names and state-dict layout (`dis_tower.{0,1,3,4,..}`,
`classifier_cls_{c}.{0,2}`); the arithmetic runs on libscan_b200.so:

  * dis_tower = [Conv3x3 + GroupNorm(32) + ReLU] x num_convs: the head_in kernels (csrc/tower.cu, csrc/gn.cu);
  * `for c in classes: Conv3x3(cat(x, act_c)) + ReLU + Conv3x3 -> weighted BCE` as two tcgen05 convolutions for all classes
    at once (block-structured weights over the two inputs, no concatenation, no Python loop over classes on the data path)
    plus the BCE kernels of csrc/cka.cu (ops.cka_class_maps_loss);
  * the gradient reversal scales with scan_scale / scan_thin_unpack.
Only CON_FUSUIN_CFG = 'concat' (what every shipped config uses) is built; 'mul' / 'mul_detached' raise.  CUDA only.
"""
import torch
from torch import nn

from . import ops
from .condgraph import _tower_conv


class GradientReversal(nn.Module):
    """layer.py:27-34."""

    def __init__(self, lambda_=1):
        super().__init__()
        self.lambda_ = lambda_

    def forward(self, x):
        return ops.grad_reverse(x, self.lambda_)


class FCOSDiscriminator_con(nn.Module):
    def __init__(self, with_GA=False, fusion_cfg="concat", num_convs=3, in_channels=256, num_classes=2, grad_reverse_lambda=-1.0,
                 grl_applied_domain="both", patch_stride=None, cfg=None):
        super().__init__()
        if in_channels != ops.C:
            raise RuntimeError("the CKA discriminator kernels are built for %d feature channels" % ops.C)
        tower = []
        for _ in range(num_convs):
            tower.append(nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1))
            tower.append(nn.GroupNorm(32, in_channels))
            tower.append(nn.ReLU())
        self.add_module("dis_tower", nn.Sequential(*tower))
        self.use_bg = False
        self.num_classes = num_classes if self.use_bg else num_classes - 1
        self.with_GA = with_GA
        self.fusion_cfg = fusion_cfg
        self.class_cond_map = []
        for i in range(self.num_classes):
            block = nn.Sequential(
                nn.Conv2d(in_channels + 1 if fusion_cfg == "concat" else in_channels, 128, kernel_size=3, stride=1, padding=1),
                nn.ReLU(),
                nn.Conv2d(128, 1, kernel_size=3, stride=1, padding=1))
            self.class_cond_map.append(block)
            self.add_module("classifier_cls_{}".format(i), block)
        self.patch_stride = patch_stride
        assert patch_stride is None or type(patch_stride) == int, "wrong format of patch stride"
        for modules in [self.dis_tower] + self.class_cond_map:
            for l in modules.modules():
                if isinstance(l, nn.Conv2d):
                    torch.nn.init.normal_(l.weight, std=0.01)
                    torch.nn.init.constant_(l.bias, 0)
        self.grad_reverse = GradientReversal(grad_reverse_lambda)
        assert grl_applied_domain == "both" or grl_applied_domain == "target"
        self.grl_applied_domain = grl_applied_domain

    def _dense_weights(self):
        """Block-structured weights of the all-classes convolutions (parameter plumbing on ~10 MB; autograd routes the dense
        gradients back to the per-class parameters)."""
        c_n = self.num_classes
        w1, b1, w2, b2 = [], [], [], []
        for c, block in enumerate(self.class_cond_map):
            wa, wb = block[0].weight, block[2].weight
            w1.append(torch.cat([wa[:, :ops.C], wa.new_zeros((128, c, 3, 3)), wa[:, ops.C:], wa.new_zeros((128, c_n - 1 - c, 3, 3))], dim=1))
            w2.append(torch.cat([wb.new_zeros((1, 128 * c, 3, 3)), wb, wb.new_zeros((1, 128 * (c_n - 1 - c), 3, 3))], dim=1))
            b1.append(block[0].bias)
            b2.append(block[2].bias)
        return torch.cat(w1, 0), torch.cat(b1, 0), torch.cat(w2, 0), torch.cat(b2, 0)

    def forward(self, feature, target, act_maps=None, domain="source"):
        assert target == 0 or target == 1 or target == 0.1 or target == 0.9
        assert domain == "source" or domain == "target"
        if not feature.is_cuda:
            raise RuntimeError("scan_b200.FCOSDiscriminator_con runs on CUDA only (no CPU fallback)")
        if self.fusion_cfg != "concat":
            raise KeyError("scan_b200 builds CON_FUSUIN_CFG = 'concat' only ('%s' is not used by any shipped config)" % self.fusion_cfg)
        if self.grl_applied_domain == "both":
            feature = self.grad_reverse(feature)
            act_maps = self.grad_reverse(act_maps)
        elif self.grl_applied_domain == "target" and domain == "target":
            feature = self.grad_reverse(feature)
        if self.patch_stride:
            raise AttributeError("'FCOSDiscriminator_con' object has no attribute 'pool'")    # the reference fails here too (:96-97)
        n, _, h, w = feature.shape
        geo = ops.Geometry([(h, w)], [1], n)
        x = ops.pack_levels(geo, [feature])
        layers = list(self.dis_tower)
        for i in range(0, len(layers), 3):
            conv, gn = layers[i], layers[i + 1]
            x, stats = _tower_conv(geo, conv.weight, None, x, gn=(conv.bias, gn.eps))
            x = ops.gn_relu_levels(geo, gn.weight, gn.bias, gn.eps, x, conv_bias=conv.bias, stats=stats)
        x_rows = ops.join_rows(geo, x)
        maps32 = ops.thin_pack_maps(geo, act_maps, 0 if self.use_bg else 1, self.num_classes)
        w1, b1, w2, b2 = self._dense_weights()
        return ops.cka_class_maps_loss(geo, x_rows, maps32, w1, b1, w2, b2, float(target), self.num_classes)
