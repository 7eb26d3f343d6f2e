#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "conv3x3" 2>&1 | tail -25 > gpurun_out/p_conv.log
tail -12 gpurun_out/p_conv.log | cut -c1-250
timeout 300 python tools/bench_conv.py 16 > gpurun_out/p_bench_conv.log 2>&1
tail -12 gpurun_out/p_bench_conv.log
