#!/bin/bash
# Development aid: build A/B variants of the library with different K2 ring geometry into gpurun_out-free paths
# (build/variants/*.so, selected at run time with PRB_LIB=...).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  IFS=: read -r chunk stages ctas <<< "$v"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
       -DPRB_K2_CHUNK=$chunk -DPRB_K2_STAGES=$stages -DPRB_K2_MIN_CTAS=$ctas \
       -o build/variants/lib_${chunk}_${stages}_${ctas}.so pyrad_b200/csrc/api.cu 2>&1 | grep -E "rror" || true
  echo built $v
done
