#!/usr/bin/env bash
# Developer tool: build variants of libnbody_b200.so that differ only in the fast-lookup force kernel's launch
# configuration (-DNB_LUTF_*), into tools/variants/, so that one GPU call can time them all:
#   NB_B200_LIB=tools/variants/libnb_<tag>.so python tools/time_modes.py 131072 int8_sim
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/../nbody_cosmological_simulation_b200/csrc"
OUT=../../tools/variants; mkdir -p "$OUT" build
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -fmad=false)
# tag:minb:unroll:ipt:threads
for v in "$@"; do
  IFS=: read -r tag minb unroll ipt threads <<< "$v"
  nvcc "${FLAGS[@]}" -DNB_LUTF_MINB=$minb -DNB_LUTF_UNROLL=$unroll -DNB_LUTF_IPT=$ipt -DNB_LUTF_THREADS=$threads -Xptxas -v -c accel.cu -o build/accel_$tag.o 2> build/accel_$tag.log &
done
wait
for v in "$@"; do
  IFS=: read -r tag _ <<< "$v"
  objs=(); for s in api integrate quantize maxdist energy metrics runtime; do objs+=(build/$s.o); done
  nvcc -shared -o "$OUT/libnb_$tag.so" build/accel_$tag.o "${objs[@]}" -gencode arch=compute_100a,code=sm_100a
  echo "$tag: $(grep -A2 'ForceF32ILi[23]ELi5E' build/accel_$tag.log | grep Used | sed 's/ptxas info    ://' | tr '\n' ' ')"
done
