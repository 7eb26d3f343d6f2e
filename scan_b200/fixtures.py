"""Deterministic "trained-like" weights for tests and the benchmark.

Random-init weights of the reference give a flat softmax; to reach the informative regimes of the
target branch (DBSCAN noise kept, >=2 clusters, border points) the fixture draws
every tensor from numpy's MT19937 stream by a rule keyed on the state-dict name, then applies
two gains: `kernel_gain` on the layer that emits the conditioned kernels and `gn_gain` on the
affine of the last head_in GroupNorm.  Works for any module exposing the reference's state-dict
(the reference itself, the oracle, and scan_b200.GRAPHModule), so all three get identical bits.
"""
import zlib

import numpy as np
import torch

_STD001 = ("middle_tower", "proto_cls", "cond_2", "cond_1", "gcn_layer")


def fixture_state_dict(module_or_state, seed=1234, kernel_gain=1.0, gn_gain=1.0):
    state = module_or_state.state_dict() if hasattr(module_or_state, "state_dict") else module_or_state
    out = {}
    for name in sorted(state.keys()):
        t = state[name]
        rs = np.random.RandomState((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31 - 1))
        shape = tuple(t.shape)
        z = rs.standard_normal(shape).astype(np.float32)
        if name == "prototype":
            v = z
        elif t.dim() == 1 and name.endswith(".weight"):      # GroupNorm / LayerNorm scale
            v = 1.0 + 0.1 * z
        elif t.dim() == 1:                                    # every bias
            v = 0.05 * z
        else:
            fan_in = int(np.prod(shape[1:]))
            std = 0.01 if any(k in name for k in _STD001) else 1.0 / np.sqrt(fan_in)
            v = std * z
        out[name] = torch.from_numpy(np.ascontiguousarray(v)).to(t.dtype)
    kernel_layer = "cond_nx1.weight" if "cond_rnn.weight_ih_l0" in out else "cond_2.weight"
    out[kernel_layer] = out[kernel_layer] * kernel_gain
    gn = sorted(k for k in out if k.startswith("head_in.middle_tower.") and out[k].dim() == 1
                and not _is_conv_bias(k, out))
    if gn:
        last = gn[-1].rsplit(".", 1)[0]
        out[last + ".weight"] = out[last + ".weight"] * gn_gain
        out[last + ".bias"] = out[last + ".bias"] * gn_gain
    return out


def _is_conv_bias(key, state):
    w = key.rsplit(".", 1)[0] + ".weight"
    return w in state and state[w].dim() == 4
