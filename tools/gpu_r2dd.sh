#!/bin/bash
mkdir -p gpurun_out
SCAN_SUSTAINED_DIAG=1 timeout 600 python bench.py --steps 20 --warmup 5 --sustained 4 > gpurun_out/dd_bench.log 2> gpurun_out/dd_bench.err
grep -n "sustained per-step\|memory allocated" gpurun_out/dd_bench.err | cut -c1-1200
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/dd_bench.log") if x.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print(d["sustained"])
print(d["e2e"]["device_ms_per_step"]); print(d["dbscan_points_per_level"])
PY
