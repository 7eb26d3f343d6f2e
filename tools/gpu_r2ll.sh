#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/ll_bench_default.log 2>gpurun_out/ll_bench_default.err; echo "default bench rc=$? wall=$(( $(date +%s) - t0 )) s"
t0=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/ll_ref_default.log 2>gpurun_out/ll_ref_default.err; echo "reference arm rc=$? wall=$(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/ll_bench_default.log") if x.startswith("{")][-1])
print(d["steps"], d["warmup"], round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d["sustained"] and round(d["sustained"]["value"],1), d["gpu_launches"])
print({k:v for k,v in d["roofline"].items() if k not in ("table","note","peak_source")})
print(d["e2e"]["device_ms_per_step"])
r=json.loads([x for x in open("gpurun_out/ll_ref_default.log") if x.startswith("{")][-1]); print({k:(v if not isinstance(v,dict) else "...") for k,v in r.items()})
PY
