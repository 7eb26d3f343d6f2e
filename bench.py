#!/usr/bin/env python
"""Benchmark of the condgraph middle head (BASELINE.json metric: fwd+bwd images/s on B200 + roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the hot path over one batch: source branch fwd+bwd on `--images` source images, then target
branch fwd+bwd on `--images` target images (synthetic FPN features of Cityscapes shape, 800x1344 padded, 22 400
locations/image, 256 channels, K = 9 classes; BASELINE.json configs[1]: 8 + 8 images per GPU).  Weights are the seeded
fixture, `--settle` untimed source steps that fill the paradigm buffer, then `fixtures.fit_trained_like`: a closed-form
fit of the manifestation layer that makes the activation maps background-dominant with confident object regions like
a trained model, so that the target-domain DBSCAN sees a realistic number of points.

  value      whole-job images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same through the public module call with HOST (pinned) inputs: H2D of the step's FPN features and
             D2H of its losses inside the timed region; the batch of step i+1 is copied on a side stream while step i
             computes (a double-buffered input pipeline), the first copy is exposed
  roofline   conditional-convolution forward kernel: algorithmic bytes / its CUDA-event duration,
             against MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the CPU oracle port of the reference (oracle/condgraph_oracle.py) on the host cores, bounded sample:
             1 source + 1 target image per step (BASELINE.json configs[0])
N > 1: one process per GPU (torchrun); images are sharded (weak scaling: --images per GPU); the only collective is the
[K, 257] prototype sum|count all-reduce of every source step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FULL_SHAPES = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
L_PER_IMAGE = sum(h * w for h, w in FULL_SHAPES)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="scan_b200", choices=["scan_b200", "reference"])
    ap.add_argument("--images", type=int, default=8, help="source images (= target images) per GPU per step")
    ap.add_argument("--settle", type=int, default=6, help="untimed source steps that fill the paradigm buffer before the fit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1, help="attention dropout (reference train-mode value 0.1)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_batches(n_images, device=None, pinned=False):
    from scan_b200.synthetic import make_workload
    src_f, src_t = make_workload(n_images, 8, seed=1234, dir_seed=77)
    tgt_f, _ = make_workload(n_images, 8, seed=4321, dir_seed=77)
    if pinned:
        src_f = [f.pin_memory() for f in src_f]
        tgt_f = [f.pin_memory() for f in tgt_f]
    return src_f, src_t, tgt_f


def pretrain(module, feats, targets, steps):
    """`steps` source forward passes on one image (they fill the 3 paradigm slots), then the closed-form trained-like fit"""
    from scan_b200.fixtures import fit_trained_like
    module.train()
    with torch.no_grad():
        for _ in range(steps):
            module(None, [f[:1] for f in feats], targets=targets[:1], mode="source")
    fit_trained_like(module, [f[:1] for f in feats], targets[:1])


def one_step(module, src, src_targets, tgt, cots, ready=None):
    """source fwd+bwd, target fwd+bwd.  Cotangents stand in for the FCOS head / discriminator gradients.
    ready: optional (event, event): the copy-stream events after which the source / target inputs are valid."""
    res = []
    for idx, (mode, feats) in enumerate((("source", src), ("target", tgt))):
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready[idx])
        feats = [f.requires_grad_(True) for f in feats]
        if mode == "source":
            out = module(None, feats, targets=src_targets, mode="source")
        else:
            out = module(None, feats, targets=None, mode="target", forward_target=True)
        out_feats, loss_graph, act_loss, acts = out
        scalars = []
        if loss_graph is not None:
            scalars += [v for v in loss_graph if torch.is_tensor(v)]
        if torch.is_tensor(act_loss):
            scalars.append(act_loss)
        # cotangents are fed straight into autograd (no extra multiply / reduce kernels in the timed region)
        torch.autograd.backward(list(out_feats) + list(acts) + scalars,
                                list(cots[0]) + list(cots[1]) + [torch.ones_like(v) for v in scalars])
        res.append(torch.stack([v.detach().float() for v in scalars]) if scalars else None)
        for f in feats:
            f.grad = None
    return res


def cpu_reference_run(args, steps, warmup):
    """The reference's CPU implementation of the path = the oracle port, on all host threads; bounded sample 1+1 images."""
    from oracle.condgraph_oracle import build_oracle
    from scan_b200.config import scan_cfg
    from scan_b200.fixtures import fixture_state_dict
    torch.set_num_threads(os.cpu_count())
    cfg = scan_cfg("c2f")
    m = build_oracle(cfg)
    m.load_state_dict(fixture_state_dict(m, seed=99))
    m.multihead_attn.p_drop = args.dropout
    src_f, src_t, tgt_f = make_batches(1)
    pretrain(m, src_f, src_t, args.settle)
    m.train()
    shapes_f = [(1, 256, h, w) for h, w in FULL_SHAPES]
    shapes_a = [(1, 9, h, w) for h, w in FULL_SHAPES]
    g = torch.Generator().manual_seed(5)
    cots = ([torch.randn(s, generator=g) / 1e5 for s in shapes_f], [torch.randn(s, generator=g) / 1e5 for s in shapes_a])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step(m, [f.clone() for f in src_f], src_t, [f.clone() for f in tgt_f], cots)
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:]) / max(steps, 1)
    return 2.0 / t, t


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        ips, t = cpu_reference_run(args, args.steps, args.warmup)
        sample = "1 source + 1 target synthetic image (800x1344, K=9) per step, oracle port of the reference on host cores"
        line = {"impl": "reference", "metric": "condgraph middle-head fwd+bwd images/s", "value": ips, "unit": "images/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "Cityscapes->Foggy VGG16 SCAN config, condgraph middle head fwd+bwd, 800x1344 FPN features, "
                                       "8 classes + bg; CPU arm runs 1+1 images per step", "settle_steps": args.settle},
                "cpu_baseline": {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
                "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from scan_b200 import _lib, ops
    from scan_b200.condgraph import build_condgraph
    from scan_b200.config import scan_cfg
    from scan_b200.fixtures import fixture_state_dict

    if os.environ.get("SCAN_CUDNN_BENCHMARK", "1") == "1":
        torch.backends.cudnn.benchmark = True     # fixed shapes: let cuDNN pick the tower-convolution algorithms by measurement
    cfg = scan_cfg("c2f")
    module = build_condgraph(cfg, 256)
    module.load_state_dict(fixture_state_dict(module, seed=99))
    module.to(dev)
    module.multihead_attn.p_drop = args.dropout
    n = args.images
    src_h, src_t, tgt_h = make_batches(n, pinned=True)
    src_d = [f.to(dev) for f in src_h]
    tgt_d = [f.to(dev) for f in tgt_h]
    pretrain(module, src_d, src_t, args.settle)
    if world > 1:
        from scan_b200 import dist as sdist
        sdist.attach(module)
        for p in module.parameters():
            dist.broadcast(p.data, 0)
        dist.broadcast(module.prototype, 0)
    module.train()
    g = torch.Generator().manual_seed(5)
    # the module returns channels-last feature tensors; the FCOS head that consumes them hands back gradients in the same
    # memory format (cuDNN's backward-data follows its input), so the stand-in cotangents are channels-last as well
    cots = ([(torch.randn((n, 256, h, w), generator=g) / 1e5).to(dev).contiguous(memory_format=torch.channels_last)
             for h, w in FULL_SHAPES],
            [(torch.randn((n, 9, h, w), generator=g) / 1e5).to(dev) for h, w in FULL_SHAPES])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        one_step(module, [f.detach() for f in src_d], src_t, [f.detach() for f in tgt_d], cots)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    profiling = os.environ.get("SCAN_PROFILE") == "1"   # ncu --profile-from-start off: capture the timed region only
    if profiling:
        torch.cuda.profiler.start()
    ops.TIMERS.clear()
    ops.TIMING["on"] = not profiling
    _lib.CALLS["n"] = 0
    _lib.CALLS["launches"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step(module, [f.detach() for f in src_d], src_t, [f.detach() for f in tgt_d], cots)
    e1.record()
    barrier()
    if profiling:
        torch.cuda.profiler.stop()
    ops.TIMING["on"] = False
    launches = _lib.CALLS["launches"]
    ms = e0.elapsed_time(e1)
    kernel_ms = ops.timers_summary()
    module.record = True      # one untimed recorded step: DBSCAN point counts and node counts of this workload
    one_step(module, [f.detach() for f in src_d], src_t, [f.detach() for f in tgt_d], cots)
    module.record = False
    dbscan_info = module.last.get("dbscan_info")
    n_nodes_t = module.last.get("sample_meta").n_nodes if module.last.get("sample_meta") is not None else 0
    module.last = {}

    # ---------------- end-to-end timing: host inputs, H2D + D2H inside ----------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d2h = 0
    # Input pipeline of a training loop: the pinned host batch of step i+1 is copied on a side stream while step i computes
    # (double buffering).  Every step's H2D copy and result read-back happen inside the timed region; the first copy is
    # fully exposed.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    bufs = [([torch.empty(f.shape, device=dev) for f in src_h], [torch.empty(f.shape, device=dev) for f in tgt_h]) for _ in range(2)]
    consumed = [None, None]   # event: the step that read buffer set b has finished

    def stage(b):
        with torch.cuda.stream(copy_stream):
            if consumed[b] is not None:
                copy_stream.wait_event(consumed[b])
            evs = []
            for dst, src_ in ((bufs[b][0], src_h), (bufs[b][1], tgt_h)):     # the source pass can start before the target batch is in
                for d_, h_ in zip(dst, src_):
                    d_.copy_(h_, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                evs.append(ev)
        return evs

    barrier()
    f0.record()
    ready = stage(0)
    for i in range(args.steps):
        b_ = i & 1
        cur = ready
        if i + 1 < args.steps:
            ready = stage(b_ ^ 1)
        res = one_step(module, [x.detach() for x in bufs[b_][0]], src_t, [x.detach() for x in bufs[b_][1]], cots, ready=cur)
        consumed[b_] = torch.cuda.Event()
        consumed[b_].record(main_stream)
        host = [r.cpu() for r in res if r is not None]
        d2h = sum(h.numel() * 4 for h in host)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    sampler.stop_flag = True
    h2d = sum(f.numel() * 4 for f in src_h) + sum(f.numel() * 4 for f in tgt_h)

    t_all = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_all[0]), float(t_all[1])
    images = 2 * n * world * args.steps
    value = images / (ms / 1e3)
    e2e = images / (ms_e2e / 1e3)

    if rank == 0:
        peak, peak_src = peaks()
        # roofline of the dominant hand-written kernel: conditional conv forward, 24.7 MB algorithmic per image
        # (rows 22400*1024 B + K maps 22400*9*4 B + labels 22400*8 B); source launches carry labels
        cc = kernel_ms.get("condconv_fwd", {"ms": 0.0, "calls": 0})
        bytes_per_launch = n * L_PER_IMAGE * (1024 + 9 * 4) + n * L_PER_IMAGE * 8 * 0.5
        ach = bytes_per_launch / (cc["ms"] / max(cc["calls"], 1) / 1e3) / 1e9 if cc["calls"] else None
        roofline = {"kernel": "condconv_fwd_ts_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": (ach / peak) if ach else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture at 8 images
                    # (profiles/r01_hot_kernels_ncu_summary.txt: 185.3 + 9.3 MB), scaled to the images of this run
                    "traffic": 194.6e6 * n / 8, "traffic_unit": "bytes/launch", "algorithmic_bytes": bytes_per_launch,
                    "peak_source": peak_src}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            ips, t = cpu_reference_run(args, 3, 1)
            cpu = {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": "3 steps of 1 source + 1 target image (800x1344, K=9), oracle port on host cores"}
        line = {"metric": "condgraph middle-head fwd+bwd images/s", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tf32 tensor-core conditional conv, fp32 accumulate)",
                "data": "synthetic",
                "config": {"workload": "Cityscapes->Foggy VGG16 SCAN config, condgraph middle head fwd+bwd, %d source + %d target "
                                       "synthetic images per GPU per step, 800x1344 FPN features, 8 classes + bg" % (n, n),
                           "parallelism": "dp%d (image shards, prototype all-reduce)" % world, "settle_steps": args.settle,
                           "attention_dropout": args.dropout,
                           "towers": "3x3 convolutions = cuDNN NHWC tf32 implicit GEMM (torch's default allow_tf32), "
                                     "cudnn.benchmark %s" % ("on" if torch.backends.cudnn.benchmark else "off"),
                           "e2e_input_pipeline": "pinned host batch of step i+1 copied on a side stream during step i",
                           "l2_note": "inputs 2x%d MB per step exceed the 126 MB L2" % (n * L_PER_IMAGE * 1024 // 2 ** 20)},
                "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
                "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in kernel_ms.items()},
                "dbscan_points_per_level": dbscan_info[:, 0].tolist() if dbscan_info is not None else None,
                "target_nodes": n_nodes_t}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
