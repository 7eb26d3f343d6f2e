#!/bin/bash
# round 2, call SS: EXPERIMENT (code removed afterwards, see DESIGN.md "Things tried that did not pay"): fused autograd nodes (pack + conv1,
# GroupNorm + conv2) with the weight gradient on a side stream: whole suite, A/B bench (SCAN_B200_OVERLAP existed only in that tree)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=900 TAILN=14 run ss_gpu_tests python -m pytest tests -q -m gpu --tb=short
TMO=300 TAILN=1 run ss_bench_overlap python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
SCAN_B200_OVERLAP=0 TMO=300 TAILN=1 run ss_bench_inorder python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=300 TAILN=1 run ss_bench_overlap2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ["ss_bench_overlap", "ss_bench_inorder", "ss_bench_overlap2"]:
    try:
        d = json.loads([x for x in open("gpurun_out/%s.log" % f) if x.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "sust", d.get("sustained") and round(d["sustained"]["value"], 1), d["clocks"])
PY
