"""Import shim for the UNMODIFIED reference middle head (TEST INFRASTRUCTURE ONLY).

This file is part of the oracle, i.e. test infrastructure: only `tests/`,
`tests/tools/make_golden.py` and the validation of `oracle/condgraph_oracle.py` may use it.
It can only work inside the build container, where `/root/reference` is mounted;
on the GPU box it raises `ReferenceUnavailable` and callers skip.

What it does:
  * registers permissive stub modules for `matplotlib`, `matplotlib.pyplot`, `ipdb`
    and the un-buildable `fcos_core._C` (THC-era sources, condgraph.py:17-18, layers/nms.py:5);
  * makes `nn.Module.to('cuda')` / `Tensor.cuda()` no-ops when no GPU is present
    (condgraph.py:170-237, loss.py:421);
  * restores the old-torch broadcast of `CosineEmbeddingLoss` targets shaped [1, K*K]
    (condgraph.py:479-480);
  * supplies an attribute-style cfg (yacs is absent) with the keys of SURVEY Appendix A.2.
Nothing from the reference is copied; it is imported where it lies.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SCAN_REFERENCE_ROOT", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


class _Permissive(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _stub(*a, **k):
            raise RuntimeError("stubbed symbol %s.%s was called" % (self.__name__, name))

        return _stub


class CfgNode(dict):
    """dict with attribute access and a yacs-like clone()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return self


def to_cfgnode(d):
    if isinstance(d, dict):
        return CfgNode({k: to_cfgnode(v) for k, v in d.items()})
    return d


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "fcos_core")):
        raise ReferenceUnavailable("reference tree not found at %s" % REFERENCE_ROOT)
    import torch
    from torch import nn

    sys.dont_write_bytecode = True
    for m in ("matplotlib", "matplotlib.pyplot", "ipdb", "fcos_core._C"):
        if m not in sys.modules:
            sys.modules[m] = _Permissive(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    if not torch.cuda.is_available():
        _orig_to = nn.Module.to

        def _to(self, *args, **kwargs):
            if args and isinstance(args[0], str) and args[0].startswith("cuda"):
                return self
            return _orig_to(self, *args, **kwargs)

        nn.Module.to = _to
        torch.Tensor.cuda = lambda self, *a, **k: self

    _orig_cel = nn.CosineEmbeddingLoss.forward

    def _cel(self, a, b, t):
        if t.dim() == 2:
            t = t[:, 0]
        return _orig_cel(self, a, b, t)

    nn.CosineEmbeddingLoss.forward = _cel
    _installed = True


def reference_modules():
    """Return (condgraph module, loss module, BoxList class) of the reference."""
    install()
    import warnings

    warnings.filterwarnings("ignore")
    from fcos_core.modeling.rpn.fcos import condgraph, loss  # noqa
    from fcos_core.structures.bounding_box import BoxList  # noqa

    return condgraph, loss, BoxList


def build_reference(cfg_dict, in_channels=256):
    condgraph, _, _ = reference_modules()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # condgraph.py:221 prints USE_RNN
        m = condgraph.GRAPHModule(to_cfgnode(cfg_dict), in_channels)
    return m
