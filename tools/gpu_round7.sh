#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=600 TAILN=3 run kernels python -m pytest tests/test_gpu_kernels.py -q --tb=short
TMO=600 TAILN=8 run mb8 python tools/bench_kernels.py 8
TMO=600 TAILN=8 run mb32 python tools/bench_kernels.py 32
TMO=900 TAILN=2 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
export SCAN_PROFILE=1
TMO=900 TAILN=2 run ncu_condconv ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:condconv_fwd -c 2 -o gpurun_out/prof_condconv2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
