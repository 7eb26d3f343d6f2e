#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "head_out_levels or conv3x3 or cka" 2>&1 | tail -30 > gpurun_out/w_kern.log
tail -12 gpurun_out/w_kern.log | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_module.py -q -m gpu 2>&1 | tail -80 > gpurun_out/w_module.log
grep -n "mismatch\|passed\|failed\|FAILED\|twin (\|Error" gpurun_out/w_module.log | cut -c1-220 | head -50
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/w_bench_n1.log 2>gpurun_out/w_bench_n1.err
tail -5 gpurun_out/w_bench_n1.err
python - <<'PY'
import json
for f in ["gpurun_out/w_bench_n1.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("host_enqueue_ms_per_step")); print(d["kernel_ms_per_step"])
PY
