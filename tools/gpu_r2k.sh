#!/bin/bash
# call K: post-processor parity first, then the whole GPU suite, then bench lines of the final tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "postprocessor" 2>&1 | tail -40 > gpurun_out/k_postproc.log
cat gpurun_out/k_postproc.log | tail -30
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/k_suite.log
tail -8 gpurun_out/k_suite.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench_n1.log 2>gpurun_out/k_bench_n1.err
tail -c 3000 gpurun_out/k_bench_n1.log
