#!/bin/bash
# Development aid: build A/B variants of the library with different K2 geometry into build/variants/*.so
# (selected at run time with PRB_LIB=...).  Each argument: chunk:stages:ctas[:consumers[:acc_smem]]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  IFS=: read -r chunk stages ctas cons accs <<< "$v"
  cons=${cons:-8}; accs=${accs:-0}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
       -DPRB_K2_CHUNK=$chunk -DPRB_K2_STAGES=$stages -DPRB_K2_MIN_CTAS=$ctas -DPRB_K2_CONSUMERS=$cons -DPRB_K2_ACC_SMEM=$accs \
       -o build/variants/lib_${chunk}_${stages}_${ctas}_${cons}_${accs}.so pyrad_b200/csrc/api.cu 2>&1 | grep -E "rror" || true
  echo built $v: $(cuobjdump -res-usage build/variants/lib_${chunk}_${stages}_${ctas}_${cons}_${accs}.so 2>/dev/null | grep -A1 "k2_line_sumILi8" | grep -o "REG:[0-9]* STACK:[0-9]*")
done
