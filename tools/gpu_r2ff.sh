#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-200; }
TMO=900 TAILN=1 run ff_bench_n$N python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5
TMO=300 TAILN=1 run ff_ref_n$N python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 1 --warmup 1
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/ff_bench_n$N.log") if x.startswith("{")][-1])
print(d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1))
print({k:v for k,v in d["roofline"].items() if k not in ("table","note","peak_source")})
PY
