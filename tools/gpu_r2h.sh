#!/bin/bash
# round 2, GPU call H: full GPU suite on the final tree, bench lines for profiles/, ncu launch list, sanitizer on the new kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=1800 TAILN=12 run r2h_all python -m pytest tests -m gpu -q --tb=short
TMO=900 TAILN=3 run r2h_bench python bench.py --steps 10 --warmup 3
TMO=600 TAILN=3 run r2h_bench_sim10k python bench.py --steps 10 --warmup 3 --config sim10k --no-cpu-baseline
TMO=600 TAILN=3 run r2h_bench_kitti python bench.py --steps 10 --warmup 3 --config kitti-eval
TMO=600 TAILN=3 run r2h_bench_ref python bench.py --impl reference --steps 3 --warmup 1
SCAN_PROFILE=1 TMO=900 TAILN=3 run r2h_ncu ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=6 run r2h_memcheck compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x --tb=line \
  -k "local_gcn or fcos_loss or transfer or rows_ or focal or graph_attention_block or node_classifier"
TMO=600 TAILN=6 run r2h_racecheck compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x --tb=line \
  -k "fcos_loss or transfer or rows_ or node_classifier"
