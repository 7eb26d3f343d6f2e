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


def box_label_maps(targets, level_shapes, strides):
    """[N, H_l, W_l] int64 maps: label of the LAST box that strictly contains the location centre, 0 = background
    (a plain inside-box rule, only used to estimate class means for `fit_trained_like`)."""
    out = []
    for (h, w), s in zip(level_shapes, strides):
        ys = (torch.arange(h, dtype=torch.float32) * s + s // 2)[:, None]
        xs = (torch.arange(w, dtype=torch.float32) * s + s // 2)[None, :]
        lab = torch.zeros((len(targets), h, w), dtype=torch.int64)
        for n, t in enumerate(targets):
            bx, lb = t.bbox.cpu(), t.get_field("labels").cpu()
            for k in range(bx.shape[0]):
                inside = (xs > bx[k, 0]) & (xs < bx[k, 2]) & (ys > bx[k, 1]) & (ys < bx[k, 3])
                lab[n][inside] = int(lb[k])
        out.append(lab)
    return out


@torch.no_grad()
def fit_trained_like(module, feats, targets, strides=(8, 16, 32, 64, 128), sharpness=1.0, bg_prior=0.97):
    """Closed-form stand-in for a trained model (USE_RNN configs): choose `cond_nx1.weight` (minimum-norm solution of a
    9 x 1536 linear system) so that the conditioned kernels emitted for the CURRENT prototype buffer equal an LDA
    classifier fitted to the class-mean head_in features of `feats`: background wins on background pixels, class c on
    its object pixels.  The constant needed by LDA rides on the all-ones direction (post-ReLU features have a stable
    positive sum).  Gives background-dominant activation maps with confident object regions -- the regime in which the
    reference switches the target branch on (SOLVER.INITIAL_AP50) -- without thousands of SGD steps."""
    dev = module.prototype.device
    hin = module.head_in([f.to(dev) for f in feats])
    k = module.prototype.shape[0]
    shapes = [tuple(f.shape[-2:]) for f in hin]
    labs = box_label_maps(targets, shapes, strides)
    x = torch.cat([f.permute(0, 2, 3, 1).reshape(-1, f.shape[1]) for f in hin]).double()
    y = torch.cat([l.reshape(-1) for l in labs]).to(dev)
    mu_all = x.mean(0)
    var = x.var(0).mean().clamp(min=1e-6)
    ones_sum = x.sum(1).mean()
    w_des = torch.zeros((k, x.shape[1]), dtype=torch.float64, device=dev)
    n_fg = max(k - 1, 1)
    for c in range(k):
        sel = y == c
        mu = x[sel].mean(0) if bool(sel.any()) else mu_all
        prior = bg_prior if c == 0 else (1.0 - bg_prior) / n_fg
        const = -(mu @ mu - mu_all @ mu_all) / (2 * var) + torch.log(torch.tensor(prior, dtype=torch.float64))
        w_des[c] = sharpness * ((mu - mu_all) / var + const / ones_sum)
    h, _ = module.cond_rnn(module.prototype.permute(2, 0, 1).contiguous())          # [P,K,512]
    hm = h.permute(1, 2, 0).reshape(k, -1).double()                                # [K, 512*P], (c,p) order
    target = (w_des - module.cond_nx1.bias.double()[None, :])                      # [K,256]
    wm = torch.linalg.pinv(hm) @ target                                            # [512*P, 256]
    module.cond_nx1.weight.copy_(wm.t().reshape(module.cond_nx1.weight.shape[0], -1, h.shape[0], 1).float())
    return w_des.float()
