#!/bin/bash
# Build a tuning variant of the library into scratch_libs/:  scripts/build_variant.sh NAME [extra nvcc flags...]
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared "$@" \
  -o scratch_libs/lib_$name.so wurm_b200/csrc/*.cu
