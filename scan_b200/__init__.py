"""scan_b200: B200-native condgraph middle head of SCAN (sm_100a kernels behind a C ABI)."""
from .config import scan_cfg, default_cfg, CfgNode  # noqa: F401
from .structures import BoxList  # noqa: F401


def build_condgraph(cfg, in_channels=256):
    from .condgraph import build_condgraph as _b
    return _b(cfg, in_channels)


def build_middle_head(cfg, in_channels=256):
    """fcos_core/modeling/rpn/rpn.py:215-218."""
    if cfg.MODEL.MIDDLE_HEAD.CONDGRAPH_ON:
        return build_condgraph(cfg, in_channels)
    raise RuntimeError("only the condgraph middle head is implemented")
