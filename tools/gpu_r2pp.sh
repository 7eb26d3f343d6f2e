#!/bin/bash
# round 2, call PP: GroupNorm backward without the y reads + GroupNorm statistics from the convolution epilogue: whole suite, A/B bench
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=900 TAILN=12 run pp_gpu_tests python -m pytest tests -q -m gpu --tb=short
TMO=300 TAILN=1 run pp_bench_fused python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
SCAN_B200_GN_STATS=0 TMO=300 TAILN=1 run pp_bench_sep python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ["pp_bench_fused", "pp_bench_sep"]:
    try:
        d = json.loads([x for x in open("gpurun_out/%s.log" % f) if x.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    k = d["kernel_ms_per_step"]
    print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "sust", d.get("sustained") and round(d["sustained"]["value"], 1), d["clocks"])
    print("   ", {n: round(v, 3) for n, v in k.items() if n.startswith("gn_") or n.startswith("conv3x3_rows")})
PY
