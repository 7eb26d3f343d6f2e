#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=900 TAILN=1 run bb_bench_n1 python bench.py --steps 20 --warmup 5
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bb_bench_n1.log") if x.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print(d["sustained"])
print(d["e2e"]["device_ms_per_step"])
PY
