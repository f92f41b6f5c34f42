#!/bin/bash
# development aid: build an alternative library into scratch_libs/<name>.so with extra nvcc flags
#   bash scripts/build_variant.sh nosplit -DAT2_SPLIT_LEAN=false
name=$1; shift
mkdir -p scratch_libs
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared \
  -diag-suppress 177 "$@" -I include -o scratch_libs/$name.so autotune-v2_b200/csrc/*.cu
