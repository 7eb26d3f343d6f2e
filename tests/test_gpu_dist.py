"""Multi-rank parity of the PRODUCT on the GPU: two ranks, each with its own image shard, run
GRAPHModule source steps with scan_b200.dist.attach(); the packed [K, 257] class sum|count buffer is all-reduced inside
update_prototype_ensemble and every rank applies the identical EMA.

Checked:  (1) the `prototype` buffer stays bit-identical across ranks;  (2) it equals the reference EMA (the CPU oracle's
update_prototype) fed with the COUNT-WEIGHTED mean of the per-shard class means;  (3) each rank's labels / node rows equal the
oracle run on that rank's shard alone (sharding changes nothing else).

Backend: NCCL when the box has >= 2 GPUs (one rank per GPU); on a single-GPU box both ranks share cuda:0 and the collective
runs over gloo (NCCL refuses two ranks on one device) -- the product code path (dist.all_reduce on the packed CUDA tensor) is
the same."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, backend, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    import harness
    from oracle.condgraph_oracle import build_oracle
    from scan_b200 import dist as sdist
    from scan_b200.condgraph import build_condgraph
    from scan_b200.fixtures import fixture_state_dict

    cfg, case, src_feats, src_targets, _ = harness.build_case("c2f_small")     # 2 images: one per rank
    n_img = src_feats[0].shape[0]
    mine = sdist.shard(list(range(n_img)), rank, world)
    feats = [f[mine] for f in src_feats]
    targets = [src_targets[i] for i in mine]

    m = build_condgraph(cfg, 256)
    m.load_state_dict(fixture_state_dict(m, seed=99))
    m.to(dev).train()
    m.multihead_attn.p_drop = 0.0
    m.record = True
    sdist.attach(m)
    proto0 = m.prototype.detach().cpu().clone()

    # reference: the oracle on this rank's shard alone (per-shard class means + counts), EMA with the global weighted mean
    o = build_oracle(cfg)
    o.load_state_dict(fixture_state_dict(o, seed=99))
    o.train()
    o.multihead_attn.p_drop = 0.0
    steps = 4          # crosses the PROTO_ITER = 3 boundary: slot fill, then the shift branch
    for step in range(steps):
        m(None, [f.to(dev) for f in feats], targets=[t.to(dev) for t in targets], mode="source")
        proto_before, counter_before = o.prototype.clone(), o.counter_rnn.counter
        o(None, [f.clone() for f in feats], targets=targets, mode="source")
        # bit-exact integer results per shard
        assert torch.equal(m.last["node_rows"].cpu(), o.last["node_rows"]) and torch.equal(m.last["node_labels"].cpu(), o.last["node_labels"])
        # undo the oracle's LOCAL EMA and redo it with the all-reduced (count-weighted) class means
        k = o.K
        lab = o.last["node_labels"]
        cnt = torch.bincount(lab, minlength=k).float()
        sums = o.last["prototype_batch"] * cnt[:, None]
        packed = torch.cat([sums, cnt[:, None]], dim=1).to(dev)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        packed = packed.cpu()
        gmean = torch.where(packed[:, -1:] > 0, packed[:, :-1] / packed[:, -1:].clamp(min=1), torch.zeros(k, sums.shape[1]))
        o.prototype.copy_(proto_before)
        o.counter_rnn.counter = counter_before
        o.update_prototype(gmean)
        got = m.prototype.detach().cpu()
        want = o.prototype
        err = float((got - want).abs().max() / want.abs().max())
        assert err <= 1e-3, "step %d: prototype differs from the count-weighted reference EMA by %.3e" % (step, err)
        mine_p = m.prototype.detach().clone() if backend == "nccl" else m.prototype.detach().cpu()   # gloo gathers host tensors
        gathered = [torch.zeros_like(mine_p) for _ in range(world)]
        dist.all_gather(gathered, mine_p)
        assert torch.equal(gathered[0], gathered[1]), "prototype diverged across ranks at step %d" % step
    assert not torch.equal(proto0, m.prototype.detach().cpu())
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_two_ranks_keep_the_paradigm_replicated_and_equal_to_the_weighted_reference_ema():
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
