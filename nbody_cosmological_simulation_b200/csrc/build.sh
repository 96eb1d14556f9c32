#!/usr/bin/env bash
# Build libnbody_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libnbody_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
       --expt-relaxed-constexpr -fmad=false -Xptxas -v)   # -fmad=false: every FMA in this library is written explicitly; nvcc otherwise contracts even __fmul2_rn+__fadd2_rn
SRCS=(api.cu integrate.cu accel.cu quantize.cu maxdist.cu energy.cu metrics.cu runtime.cu galaxy_init.cu)
OBJS=()
mkdir -p "${HERE}/build"
pids=()
for s in "${SRCS[@]}"; do
  o="${HERE}/build/${s%.cu}.o"
  OBJS+=("$o")
  ( "$NVCC" "${FLAGS[@]}" -c "${HERE}/${s}" -o "$o" > "${HERE}/build/${s%.cu}.log" 2>&1 ) &
  pids+=($!)
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then fail=1; echo "nvcc failed for ${SRCS[$i]}:"; cat "${HERE}/build/${SRCS[$i]%.cu}.log"; fi
done
[ "$fail" = 0 ] || exit 1
"$NVCC" -shared -o "$OUT" "${OBJS[@]}" -gencode arch=compute_100a,code=sm_100a
echo "built $OUT"
