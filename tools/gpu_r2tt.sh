#!/bin/bash
# round 2, call TT: the final tree once more: whole GPU suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=900 TAILN=4 run tt_gpu_tests python -m pytest tests -x -q -m gpu
TMO=300 TAILN=2 run tt_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=600 TAILN=1 run tt_bench_default python bench.py
TMO=600 TAILN=1 run tt_bench_n1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ["tt_bench_default", "tt_bench_n1"]:
    d=json.loads([x for x in open("gpurun_out/%s.log"%f) if x.startswith("{")][-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1), d["gpu_launches"], d["steps"], d["warmup"])
    print({k:v for k,v in d["roofline"].items() if k not in ("table","note")})
PY
