#!/bin/bash
# ncu --set full of the tower kernels INSIDE one bench step (8 images per launch): duration, DRAM bytes, tensor pipe
mkdir -p gpurun_out
SCAN_PROFILE=1 timeout 1200 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name 'regex:conv3x3_kernel|conv_wgrad_kernel' --launch-count 24 -o gpurun_out/kk_conv_step python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline --sustained 0 > gpurun_out/kk_ncu.log 2>&1
ncu -i gpurun_out/kk_conv_step.ncu-rep --page raw --csv > gpurun_out/kk_conv_step_raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/kk_conv_step_raw.csv > gpurun_out/kk_conv_step_pick.txt 2>/dev/null
grep -c "^----" gpurun_out/kk_conv_step_pick.txt; head -40 gpurun_out/kk_conv_step_pick.txt
