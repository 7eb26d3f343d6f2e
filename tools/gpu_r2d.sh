#!/bin/bash
# round 2, GPU call D: new-kernel tests, ncu --set full of the hot kernels inside the bench step, compute-sanitizer
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=600 TAILN=30 run r2d_kernels python -m pytest tests/test_gpu_kernels.py -q --tb=short
TMO=1200 TAILN=12 run r2d_module python -m pytest tests/test_gpu_module.py -q --tb=short
TMO=600 TAILN=3 run r2d_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline
SCAN_PROFILE=1 TMO=1500 TAILN=3 run r2d_ncu_full ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name 'regex:attn_fwd1|attn_bwd_dq|attn_bwd_dkv|db_adj_tc|condconv_fwd_ts|condconv_bwd_rows' --launch-count 15 \
  -o gpurun_out/r02_hot -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline
ls -la gpurun_out/r02_hot.ncu-rep
TMO=900 TAILN=8 run r2d_memcheck compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x --tb=line \
  -k "gather or class_sums or focal or rows_ or transfer or node_classifier or ensemble or fcos_assign or target_sampling or pack_unpack or class_means"
TMO=600 TAILN=8 run r2d_racecheck compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x --tb=line \
  -k "class_sums or rows_ or transfer or focal or gather"
