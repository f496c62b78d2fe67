#!/bin/bash
# Build libpyrad_b200.so in-tree for sm_100a (the only target).  __graft_entry__.build() calls this.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -shared ${PRB_NVCC_EXTRA} \
     -o pyrad_b200/libpyrad_b200.so pyrad_b200/csrc/api.cu
