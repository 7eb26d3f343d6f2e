#!/bin/bash
mkdir -p gpurun_out
SCAN_SUSTAINED_DIAG=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --sustained 3 > gpurun_out/cc_bench.log 2> gpurun_out/cc_bench.err
grep -n "sustained per-step\|memory allocated" gpurun_out/cc_bench.err | cut -c1-1500
SCAN_SUSTAINED_DIAG=1 SCAN_B200_TOWERS=cudnn timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --sustained 3 > gpurun_out/cc_bench_cudnn.log 2> gpurun_out/cc_bench_cudnn.err
grep -n "sustained per-step\|memory allocated" gpurun_out/cc_bench_cudnn.err | cut -c1-1500
