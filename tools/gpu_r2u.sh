#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv3x3" 2>&1 | tail -5 > gpurun_out/u_conv.log
tail -3 gpurun_out/u_conv.log | cut -c1-250
timeout 1500 python -m pytest tests/test_gpu_module.py -q -m gpu 2>&1 | tail -80 > gpurun_out/u_module.log
grep -n "mismatch\|passed\|failed\|FAILED\|twin (" gpurun_out/u_module.log | cut -c1-220 | head -50
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/u_bench_n1.log 2>gpurun_out/u_bench_n1.err
tail -5 gpurun_out/u_bench_n1.err
python - <<'PY'
import json
for f in ["gpurun_out/u_bench_n1.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("host_enqueue_ms_per_step")); print({k:v for k,v in d["e2e"].items() if k!="result_read"}); print({k:v for k,v in d["roofline"].items() if k!="table"})
PY
