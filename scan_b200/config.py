"""Config tree for the condgraph middle head.

Mirrors the reference's yacs keys verbatim (fcos_core/config/defaults.py:619-712 for
MODEL.MIDDLE_HEAD.*, :338-351 for MODEL.FCOS.*, :691 for TEST.MODE) so a yacs `CfgNode`
coming from the reference's tools can be passed to `build_condgraph` unchanged, and so
that this package also works without yacs (it is absent from the image): `CfgNode`
below is a dict with attribute access and `clone()`, the only two things
condgraph.py:138 / loss.py:247 use.
"""
import copy


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        return CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _node(d):
    if isinstance(d, dict):
        return CfgNode({k: _node(v) for k, v in d.items()})
    return d


# defaults.py values (NOT the yaml values)
_DEFAULTS = {
    "MODEL": {
        "DEBUG_CFG": None,
        "FCOS": {
            "NUM_CLASSES": 81,
            "FPN_STRIDES": [8, 16, 32, 64, 128],
            "LOSS_ALPHA": 0.25,
            "LOSS_GAMMA": 2.0,
            "INFERENCE_TH": 0.05,
            "PRE_NMS_TOP_N": 1000,
            "NMS_TH": 0.6,
            "NUM_CONVS": 4,
        },
        "MIDDLE_HEAD": {
            "CONDGRAPH_ON": False,
            "NUM_CONVS_IN": 1,
            "NUM_CONVS_OUT": 1,
            "GCN1_OUT_CHANNEL": 256,
            "GCN2_OUT_CHANNEL": 256,
            "GCN_EDGE_PROJECT": 128,
            "GCN_EDGE_NORM": "softmax",
            "GCN_OUT_ACTIVATION": "relu",
            "CAT_ACT_MAP": True,
            "GCN_SHORTCUT": False,
            "RETURN_ACT_LOGITS": False,
            "COND_WITH_BIAS": False,
            "PROTO_WITH_BG": True,
            "ACT_LOSS": None,
            "ACT_LOSS_WEIGHT": 1.0,
            "GCN_LOSS_WEIGHT": 1.0,
            "CON_LOSS_WEIGHT": 1.0,
            "GCN_LOSS_WEIGHT_TG": 1.0,
            "PROTO_MOMENTUM": 0.95,
            "PROTO_CHANNEL": 256,
            "CON_TG_CFG": "KLdiv",
            "TRANSFER_CFG": (None,),
            "PROTO_MEAN_VAR": False,
            "IN_NORM": "GN",
            "GLOBAL_GCN": False,
            "COSINE_UPDATE_ON": False,
            "PROTO_ALIGN": False,
            "PROTO_ITER": 1,
            "USE_RNN": None,
            "GCN_SELF_TRAINING": False,
            "COND_HIDDEN_CHANNEL": 512,
            "TARGET_SAMPLING_CFG": "score_threshold",
            "DBSCAN_EPS": 3,
            "DBSCAN_THR": 0.05,
        },
    },
    "SOLVER": {"MIDDLE_HEAD": {"PLABEL_TH": (0.9,)}},
    "TEST": {"MODE": "common", "DETECTIONS_PER_IMG": 100},
}

# configs/scan/scan_vgg16_cityscapace_to_foggy.yaml:11-56, :68, :109
_SCAN_YAML = {
    "CONDGRAPH_ON": True,
    "NUM_CONVS_IN": 2,
    "NUM_CONVS_OUT": 1,
    "CAT_ACT_MAP": True,
    "IN_NORM": "GN",
    "COSINE_UPDATE_ON": True,
    "PROTO_ALIGN": True,
    "PROTO_MEAN_VAR": False,
    "PROTO_ITER": 3,
    "USE_RNN": "RNN",
    "PROTO_WITH_BG": True,
    "COND_WITH_BIAS": False,
    "PROTO_CHANNEL": 256,
    "PROTO_MOMENTUM": 0.95,
    "TRANSFER_CFG": ("NODES", "ADJ"),
    "GCN_SELF_TRAINING": False,
    "TARGET_SAMPLING_CFG": "dbscan",
    "DBSCAN_EPS": 3,
    "DBSCAN_THR": 0.05,
    "CON_TG_CFG": "KLdiv",
    "ACT_LOSS": "softmaxFL",
    "ACT_LOSS_WEIGHT": 1.0,
    "GCN_LOSS_WEIGHT": 1.0,
    "CON_LOSS_WEIGHT": 1.0,
    "GCN_LOSS_WEIGHT_TG": 1.0,
    "GLOBAL_GCN": True,
    "GCN_OUT_ACTIVATION": "relu",
    "GCN_EDGE_NORM": "cosine_detached",
    "GCN_SHORTCUT": False,
    "GCN1_OUT_CHANNEL": 256,
    "GCN2_OUT_CHANNEL": 256,
    "GCN_EDGE_PROJECT": 256,
}


def default_cfg():
    return _node(copy.deepcopy(_DEFAULTS))


def scan_cfg(name="c2f", **middle_head_overrides):
    """The three shipped SCAN configs (configs/scan/*.yaml).

    name: 'c2f'    Cityscapes->Foggy, NUM_CLASSES 9, TRANSFER_CFG ('NODES','ADJ'), TEST.MODE 'precision'
          'sim10k' Sim10k->Cityscapes, NUM_CLASSES 2, TRANSFER_CFG commented out -> (None,), TEST.MODE 'common'
          'kitti'  KITTI->Cityscapes, same as sim10k
    Extra keyword arguments override MODEL.MIDDLE_HEAD keys.
    """
    cfg = default_cfg()
    cfg.MODEL.MIDDLE_HEAD.update(copy.deepcopy(_SCAN_YAML))
    if name == "c2f":
        cfg.MODEL.FCOS.NUM_CLASSES = 9
        cfg.TEST.MODE = "precision"
    elif name in ("sim10k", "kitti"):
        cfg.MODEL.FCOS.NUM_CLASSES = 2
        cfg.MODEL.MIDDLE_HEAD.TRANSFER_CFG = (None,)
        cfg.TEST.MODE = "common"
    else:
        raise KeyError("unknown SCAN config %r" % (name,))
    for k, v in middle_head_overrides.items():
        if k not in cfg.MODEL.MIDDLE_HEAD:
            raise KeyError("unknown MODEL.MIDDLE_HEAD key %r" % (k,))
        cfg.MODEL.MIDDLE_HEAD[k] = v
    return cfg


def to_plain(cfg):
    """Deep-convert to plain dicts (used to hand the same config to the oracle shim)."""
    if isinstance(cfg, dict):
        return {k: to_plain(v) for k, v in cfg.items()}
    return cfg


# FPN level geometry of the headline workload: 800x1333 padded to 800x1344 (SIZE_DIVISIBILITY 32)
CITYSCAPES_LEVEL_SHAPES = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
