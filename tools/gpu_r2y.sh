#!/bin/bash
# call Y (2 GPUs): NCCL prototype test, 2-rank bench (async all-reduce + early head_out feature half), 1-rank bench on the same box
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log | cut -c1-400; }
TMO=600 TAILN=5 run y_dist python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short
TMO=900 TAILN=2 run y_bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5
TMO=600 TAILN=2 run y_bench1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ["gpurun_out/y_bench2.log","gpurun_out/y_bench1.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
