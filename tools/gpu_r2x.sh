#!/bin/bash
mkdir -p gpurun_out
SCAN_E2E_NOSTAGE=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/x_bench_nostage.log 2>gpurun_out/x_bench_nostage.err
python - <<'PY'
import json
for f in ["gpurun_out/x_bench_nostage.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d["e2e"]["device_ms_per_step"]); print(d["e2e"]["host_ms_per_step"])
PY
tail -3 gpurun_out/x_bench_nostage.err
