#!/bin/bash
# call Q: own tower convolutions in the module — whole GPU suite, then bench with both tower implementations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/q_suite.log
tail -30 gpurun_out/q_suite.log | cut -c1-260
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/q_bench_n1.log 2>gpurun_out/q_bench_n1.err
python - <<'PY'
import json
for f in ["gpurun_out/q_bench_n1.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("host_enqueue_ms_per_step")); print(d.get("kernel_ms_per_step"))
PY
SCAN_B200_TOWERS=cudnn timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/q_bench_n1_cudnn.log 2>gpurun_out/q_bench_n1_cudnn.err
python - <<'PY'
import json
for f in ["gpurun_out/q_bench_n1_cudnn.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
