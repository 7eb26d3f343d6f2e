#!/bin/bash
# evidence run of the final tree: whole GPU suite, smoke, all bench lines, launch list (profiles/r02_* are copied from its output)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-200; }
TMO=1800 TAILN=3 run fin_gpu_tests python -m pytest tests -m gpu -q --tb=short
TMO=300 TAILN=2 run fin_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 TAILN=1 run fin_bench_n1 python bench.py --steps 20 --warmup 5
TMO=900 TAILN=1 run fin_bench_sim10k python bench.py --config sim10k --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run fin_bench_kitti_eval python bench.py --config kitti-eval --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run fin_bench_reference_arm python bench.py --impl reference --steps 2 --warmup 1
SCAN_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --sustained 0 > gpurun_out/fin_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/fin_launches.csv 70 > gpurun_out/fin_launches_step_n8.txt; head -12 gpurun_out/fin_launches_step_n8.txt; tail -1 gpurun_out/fin_launches_step_n8.txt
python - <<'PY'
import json
for f in ["fin_bench_n1","fin_bench_sim10k","fin_bench_kitti_eval"]:
    d=json.loads([x for x in open("gpurun_out/%s.log"%f) if x.startswith("{")][-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1), d.get("dbscan_points_per_level"), d.get("light_mode") and round(d["light_mode"]["value"],1))
    if f=="fin_bench_n1":
        print({k:v for k,v in d["roofline"].items() if k not in ("table","note")})
        for r in d["roofline"]["table"][:12]: print("  ", r["entry"], round(r["ms_per_step"],3), round(r["achieved"],1), r["unit"], round(r["frac"],3))
        print(d["eager_gpu_baseline"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"], d["source_nodes"], d["target_nodes"])
PY
