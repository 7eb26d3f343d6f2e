#!/bin/bash
# call Z: evidence run of the tree — whole GPU suite, smoke, all bench lines, launch list, sanitizers over the new kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=1800 TAILN=4 run z_gpu_tests python -m pytest tests -m gpu -q --tb=short
TMO=300 TAILN=2 run z_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 TAILN=1 run z_bench_n1 python bench.py --steps 20 --warmup 5
TMO=900 TAILN=1 run z_bench_sim10k python bench.py --config sim10k --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run z_bench_kitti_eval python bench.py --config kitti-eval --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=1 run z_bench_reference_arm python bench.py --impl reference --steps 2 --warmup 1
SCAN_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/z_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/z_launches.csv 60 > gpurun_out/z_launches_step_n8.txt; head -25 gpurun_out/z_launches_step_n8.txt; tail -1 gpurun_out/z_launches_step_n8.txt
TMO=1200 TAILN=6 run z_memcheck compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q --tb=line -k "conv3x3 or head_out_levels or cka or postprocessor"
TMO=1200 TAILN=6 run z_racecheck compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q --tb=line -k "conv3x3_rows_matches_fp64_conv and shapes2 or conv3x3_wgrad_matches_fp64 and shapes2 or postprocessor_matches_reference_golden"
