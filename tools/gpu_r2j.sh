#!/bin/bash
# round 2, GPU call J: kernel-read upload of the targets, e2e check, sanitizer over the whole kernel test file
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=1800 TAILN=8 run r2j_all python -m pytest tests -m gpu -q --tb=short
TMO=900 TAILN=3 run r2j_bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
TMO=1200 TAILN=6 run r2j_memcheck compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q --tb=line \
  -k "not dbscan and not attention and not condconv and not manifest"
TMO=900 TAILN=6 run r2j_memcheck_tc compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q --tb=line \
  -k "attention_matches_reference_view_semantics or condconv_forward_backward or dbscan_threshold_band"
