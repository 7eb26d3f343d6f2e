"""Host-side logic of the multi-GPU path with world_size-2 gloo on CPU: image sharding and the packed
[K, 257] prototype sum|count all-reduce that keeps the paradigm replicated."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scan_b200 import dist as sdist
    from oracle.condgraph_oracle import build_oracle
    from scan_b200.config import scan_cfg

    rs = np.random.RandomState(0)
    images = list(range(8))
    mine = sdist.shard(images, rank, world)
    assert mine == images[rank * 4:(rank + 1) * 4]
    # per-image node sets (identical on every rank), each rank reduces only its own images
    k, c = 9, 256
    nodes = [torch.from_numpy(rs.standard_normal((50 + 7 * i, c)).astype(np.float32)) for i in images]
    labels = [torch.from_numpy(rs.randint(0, k - 1, 50 + 7 * i)) for i in images]   # class 8 never appears
    packed = torch.zeros(k, c + 1)
    for i in mine:
        packed[:, :c].index_add_(0, labels[i], nodes[i])
        packed[:, c] += torch.bincount(labels[i], minlength=k).float()

    class Dummy(object):
        dist_group = None
    m = sdist.attach(Dummy())
    assert m.dist_group is not None
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=m.dist_group)
    all_nodes, all_labels = torch.cat(nodes), torch.cat(labels)
    want = torch.zeros(k, c + 1)
    want[:, :c].index_add_(0, all_labels, all_nodes)
    want[:, c] = torch.bincount(all_labels, minlength=k).float()
    assert torch.allclose(packed, want, rtol=1e-5, atol=1e-4)
    # every rank applies the identical EMA -> replicated prototype (global count > 0 decides `exist`)
    cfg = scan_cfg("c2f")
    torch.manual_seed(1)
    orc = build_oracle(cfg)
    batch = torch.where(packed[:, c:] > 0, packed[:, :c] / packed[:, c:].clamp(min=1), torch.zeros(k, c))
    orc.update_prototype(batch)
    gathered = [torch.zeros_like(orc.prototype) for _ in range(world)]
    dist.all_gather(gathered, orc.prototype)
    assert torch.equal(gathered[0], gathered[1])
    assert float(packed[8, c]) == 0.0 and torch.equal(orc.prototype[8], gathered[0][8])
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_prototype_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
