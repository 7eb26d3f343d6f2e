#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=1200 TAILN=4 run aa_racecheck compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q --tb=line -k "conv3x3_rows_matches_fp64_conv and shapes2 or conv3x3_wgrad_matches_fp64 and shapes2 or postprocessor_matches_reference_golden or head_out_levels"
TMO=900 TAILN=1 run aa_bench_n1 python bench.py --steps 20 --warmup 5
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/aa_bench_n1.log") if x.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print({k:v for k,v in d["roofline"].items() if k!="table"})
for r in d["roofline"]["table"][:10]: print(r["entry"], round(r["ms_per_step"],3), round(r["achieved"],1), r["unit"], round(r["frac"],3))
print(d["e2e"]["device_ms_per_step"])
PY
