#!/bin/bash
# round 2, call VV (4 GPUs): bench line of the final tree on four ranks
mkdir -p gpurun_out
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/vv_bench_n4.log 2>&1; echo "rc=$?"
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/vv_bench_n4.log") if x.startswith("{")][-1])
print(d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1), d["e2e"].get("h2d_alone_ms_per_step"))
PY
