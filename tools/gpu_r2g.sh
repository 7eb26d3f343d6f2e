#!/bin/bash
# round 2, GPU call G (2 GPUs): NCCL dist test + 2-rank bench
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
nvidia-smi -L
TMO=600 TAILN=15 run r2g_dist python -m pytest tests/test_gpu_dist.py tests/test_gpu_kernels.py -q --tb=short -k "two_ranks or local_gcn"
TMO=900 TAILN=3 run r2g_bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3
TMO=600 TAILN=3 run r2g_bench1 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
