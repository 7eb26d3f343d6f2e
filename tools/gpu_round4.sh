#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=600 TAILN=6 run kernels python -m pytest tests/test_gpu_kernels.py -q --tb=short
TMO=900 TAILN=14 run module_tc python -m pytest tests/test_gpu_module.py -q --tb=short
TMO=900 TAILN=3 run bench python bench.py --steps 10 --warmup 3
export SCAN_PROFILE=1
TMO=900 TAILN=2 run ncu_launches ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline
TMO=900 TAILN=2 run ncu_condconv ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:condconv_ -c 3 -o gpurun_out/prof_condconv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
TMO=900 TAILN=2 run ncu_dbscan ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:db_adj_tc -c 1 -o gpurun_out/prof_dbscan -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
ls -la gpurun_out | head -20
