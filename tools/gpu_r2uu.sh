#!/bin/bash
# round 2, call UU (2 GPUs): the final tree on two ranks: the 2-rank product test over NCCL and the bench line
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-200; }
TMO=300 TAILN=3 run uu_dist_test python -m pytest tests/test_gpu_dist.py -q -m gpu
TMO=600 TAILN=1 run uu_bench_n2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/uu_bench_n2.log") if x.startswith("{")][-1])
print(d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1))
PY
