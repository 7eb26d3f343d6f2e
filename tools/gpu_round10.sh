#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=900 TAILN=1 run bench_full python bench.py
TMO=600 TAILN=1 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
export SCAN_PROFILE=1
TMO=900 TAILN=1 run ncu_step ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline
