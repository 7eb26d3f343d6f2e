#!/bin/bash
# round 2, GPU call C: single-pass attention forward + focal/ensemble kernels + new bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=300 TAILN=40 run r2c_attn python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q --tb=short -k "attention or focal or ensemble"
TMO=600 TAILN=15 run r2c_kernels python -m pytest tests/test_gpu_kernels.py -q --tb=short
TMO=1200 TAILN=30 run r2c_module python -m pytest tests/test_gpu_module.py -q --tb=short
TMO=900 TAILN=3 run r2c_bench python bench.py --steps 5 --warmup 3
SCAN_PROFILE=1 TMO=900 TAILN=3 run r2c_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline
