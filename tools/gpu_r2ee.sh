#!/bin/bash
# call EE: bench lines of the final bench.py (stationary workload) + launch list + racecheck record
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-200; }
TMO=900 TAILN=1 run ee_bench_n1 python bench.py --steps 20 --warmup 5
TMO=900 TAILN=1 run ee_bench_sim10k python bench.py --config sim10k --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run ee_bench_kitti_eval python bench.py --config kitti-eval --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run ee_bench_reference_arm python bench.py --impl reference --steps 2 --warmup 1
SCAN_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ee_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --sustained 0 > gpurun_out/ee_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/ee_launches.csv 60 > gpurun_out/ee_launches_step_n8.txt; head -12 gpurun_out/ee_launches_step_n8.txt; tail -1 gpurun_out/ee_launches_step_n8.txt
python - <<'PY'
import json
for f in ["ee_bench_n1","ee_bench_sim10k","ee_bench_kitti_eval"]:
    d=json.loads([x for x in open("gpurun_out/%s.log"%f) if x.startswith("{")][-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1), d.get("dbscan_points_per_level"))
    if f=="ee_bench_n1":
        print({k:v for k,v in d["roofline"].items() if k not in ("table","note")})
        for r in d["roofline"]["table"][:12]: print("  ", r["entry"], round(r["ms_per_step"],3), round(r["achieved"],1), r["unit"], round(r["frac"],3))
        print(d["eager_gpu_baseline"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"])
PY
