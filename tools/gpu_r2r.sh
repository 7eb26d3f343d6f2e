#!/bin/bash
# call R: precise-mode segments + side-stream sampling: conv tests, module parity, bench + host profile
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv3x3" 2>&1 | tail -25 > gpurun_out/r_conv.log
tail -5 gpurun_out/r_conv.log | cut -c1-250
timeout 1500 python -m pytest tests/test_gpu_module.py tests/test_gpu_fullsize.py tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -80 > gpurun_out/r_module.log
grep -n "float mismatch\|passed\|failed\|FAILED" gpurun_out/r_module.log | cut -c1-220 | head -50
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r_bench_n1.log 2>gpurun_out/r_bench_n1.err
SCAN_HOST_PROFILE=1 timeout 400 python bench.py --steps 10 --warmup 5 > gpurun_out/r_bench_prof.log 2>gpurun_out/r_bench_prof.err
python - <<'PY'
import json
for f in ["gpurun_out/r_bench_n1.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("host_enqueue_ms_per_step")); print(d["e2e"].get("host_ms_per_step"))
PY
head -70 gpurun_out/r_bench_prof.err | cut -c1-200
