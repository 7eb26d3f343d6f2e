#!/bin/bash
# round 2, call RR: evidence of the tree with the GroupNorm changes: memcheck of the new code, ncu --set full of the GroupNorm / convolution
# kernels inside the bench step, all bench lines, the launch list
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=400 TAILN=4 run rr_memcheck compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "epilogue_group_norm or gn_relu_levels or add_relu"
SCAN_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name 'regex:gn_bwd_reduce|gn_bwd_apply|gn_apply|conv3x3_kernel|conv_gn_finalize' --launch-count 30 -o gpurun_out/rr_gn_step python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline --sustained 0 > gpurun_out/rr_ncu.log 2>&1
ncu -i gpurun_out/rr_gn_step.ncu-rep --page raw --csv > gpurun_out/rr_gn_step_raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/rr_gn_step_raw.csv > gpurun_out/rr_gn_step_pick.txt 2>/dev/null
rm -f gpurun_out/rr_gn_step.ncu-rep
TMO=900 TAILN=1 run rr_bench_n1 python bench.py --steps 20 --warmup 5
TMO=900 TAILN=1 run rr_bench_sim10k python bench.py --config sim10k --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run rr_bench_kitti_eval python bench.py --config kitti-eval --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run rr_bench_reference_arm python bench.py --impl reference --steps 2 --warmup 1
SCAN_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rr_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --sustained 0 > gpurun_out/rr_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/rr_launches.csv 70 > gpurun_out/rr_launches_step_n8.txt; head -14 gpurun_out/rr_launches_step_n8.txt; tail -1 gpurun_out/rr_launches_step_n8.txt
python - <<'PY'
import json
for f in ["rr_bench_n1","rr_bench_sim10k","rr_bench_kitti_eval"]:
    try:
        d=json.loads([x for x in open("gpurun_out/%s.log"%f) if x.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1), d.get("dbscan_points_per_level"), d.get("light_mode") and round(d["light_mode"]["value"],1))
    if f=="rr_bench_n1":
        print({k:v for k,v in d["roofline"].items() if k not in ("table","note")})
        for r in d["roofline"]["table"][:14]: print("  ", r["entry"], round(r["ms_per_step"],3), round(r["achieved"],1), r["unit"], round(r["frac"],3))
        print(d["eager_gpu_baseline"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"], d["source_nodes"], d["target_nodes"])
PY
grep -c "^----" gpurun_out/rr_gn_step_pick.txt
