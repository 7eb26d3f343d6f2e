#!/bin/bash
# call N: ncu of the tower conv kernels + wgrad scaling with problem size
mkdir -p gpurun_out
for n in 1 2 4 16; do echo "== n_images $n"; timeout 200 python tools/bench_conv.py $n 2>&1 | grep -E "tower wgrad|cta_group 2: .* TFLOP/s$|cuDNN wgrad"; done > gpurun_out/n_sizes.log 2>&1
cat gpurun_out/n_sizes.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_kernel|conv3x3_kernel|conv_wgrad_reduce" -c 6 -o gpurun_out/n_conv python tools/bench_conv.py 16 > gpurun_out/n_ncu.log 2>&1
ncu -i gpurun_out/n_conv.ncu-rep --page raw --csv > gpurun_out/n_conv_raw.csv 2>/dev/null
ls -la gpurun_out/n_conv* | head
