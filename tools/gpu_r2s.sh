#!/bin/bash
# call S: module parity with both twins, bench, launch list of one step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv3x3" 2>&1 | tail -5 > gpurun_out/s_conv.log
tail -3 gpurun_out/s_conv.log | cut -c1-250
timeout 1500 python -m pytest tests/test_gpu_module.py -q -m gpu 2>&1 | tail -80 > gpurun_out/s_module.log
grep -n "mismatch\|passed\|failed\|FAILED\|twin (" gpurun_out/s_module.log | cut -c1-220 | head -50
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench_n1.log 2>gpurun_out/s_bench_n1.err
python - <<'PY'
import json
for f in ["gpurun_out/s_bench_n1.log"]:
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print("no json", f); continue
    d=json.loads(l[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("host_enqueue_ms_per_step")); print(d["e2e"].get("host_ms_per_step")); print(d["kernel_ms_per_step"])
PY
SCAN_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/s_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/s_launches.csv 45 > gpurun_out/s_launches_step_n8.txt; head -50 gpurun_out/s_launches_step_n8.txt
