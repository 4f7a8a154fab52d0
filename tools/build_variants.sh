#!/bin/bash
# Development aid: builds variants/libpt_<name>.so from "name:-Dflag -Dflag" arguments (timed by tools/variants.py).
# usage: tools/build_variants.sh "base:-DPT_FILE_ONESHOT=0 -DPT_RCP_APPROX=0" "oneshot:-DPT_RCP_APPROX=0" ...
set -e
cd "$(dirname "$0")/../cpppathtracer_b200/csrc"
mkdir -p ../../variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ \
       -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -Xptxas -v -I../../include --shared --fmad=false $flags \
       -o ../../variants/libpt_${name}.so pt_api.cu pt_png.cpp -lz 2> ../../variants/build_${name}.log &
done
wait
for spec in "$@"; do
  name="${spec%%:*}"
  echo "== $name"; grep -A3 "Compiling entry function ._ZN2pt8k_bounceILi[01]ELb0" ../../variants/build_${name}.log | grep -E "Used|error" | sed 's/^ptxas info    : //'
done
