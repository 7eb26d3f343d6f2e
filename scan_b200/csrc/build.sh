#!/bin/bash
# Build libscan_b200.so (sm_100a only) in-tree: scan_b200/libscan_b200.so
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libscan_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -fmad=true"
OBJS=()
mkdir -p "$HERE/build"
pids=()
for f in core layout gn manifest assign losses proto condconv gemm transfer rowsmlp fcosloss postproc tower cka attention attention_t5 attention_t5_bwd dbscan dbscan_tc; do
  "$NVCC" $FLAGS ${SCAN_PTXAS_V:+-Xptxas -v} -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
  pids+=($!)
  OBJS+=("$HERE/build/$f.o")
done
for p in "${pids[@]}"; do wait $p; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "${OBJS[@]}" -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
