#!/bin/bash
# round 2, GPU call E: double-buffered attention backward, GCN + FCOS-loss kernels, all bench configs
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=300 TAILN=25 run r2e_attn python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q --tb=short -k "attention"
TMO=600 TAILN=25 run r2e_kernels python -m pytest tests/test_gpu_kernels.py -q --tb=short
TMO=1200 TAILN=12 run r2e_module python -m pytest tests/test_gpu_module.py -q --tb=short
TMO=900 TAILN=3 run r2e_bench python bench.py --steps 10 --warmup 3
TMO=600 TAILN=3 run r2e_bench_sim10k python bench.py --steps 5 --warmup 3 --config sim10k --no-cpu-baseline
TMO=600 TAILN=3 run r2e_bench_kitti python bench.py --steps 10 --warmup 3 --config kitti-eval --no-cpu-baseline
SCAN_PROFILE=1 TMO=900 TAILN=3 run r2e_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline
