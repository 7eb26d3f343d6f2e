"""Multi-GPU plumbing: one process per GPU, images sharded by rank, and ONE collective on the data path --
the all-reduce (sum) of the packed per-class [K, 257] prototype sum|count buffer of every source step, after which every
rank applies the identical paradigm EMA (scan_proto_update), so the `prototype` buffer stays replicated.
The reference has no such collective: its DDP uses broadcast_buffers=False (tools/train_net_da.py:427-432) and lets the
buffer diverge; the north-star adds it.  ~9 KB per step: latency-bound, NCCL over NVLink is sufficient.
"""
import torch.distributed as dist


def attach(module, group=None):
    """Make `module.update_prototype_ensemble` all-reduce the class sums over `group` (default: the world)."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    module.dist_group = group if group is not None else dist.group.WORLD
    return module


def shard(items, rank=None, world=None):
    """Contiguous block of `items` (images) owned by `rank`."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per = (len(items) + world - 1) // world
    return items[rank * per:(rank + 1) * per]
