#!/bin/bash
# Build libsyntalker_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libsyntalker_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
SRCS="st_api.cu st_kernels.cu st_gemm_tc.cu"
newest=$(ls -t $SRCS st_internal.cuh ../../include/syntalker_b200.h build.sh | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "up to date: $OUT"; exit 0
fi
objs=""
pids=""
for s in $SRCS; do
  o="${s%.cu}.o"
  $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -c "$s" -o "$o" &
  pids="$pids $!"
  objs="$objs $o"
done
for p in $pids; do wait $p; done
$NVCC -shared -o "$OUT" $objs
echo "built $OUT"
