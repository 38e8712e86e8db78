#!/bin/bash
# Builds experiment variants of libunigen_b200.so (different -D switches of gemm_tc.cu) into build/exp/.
set -e
cd "$(dirname "$0")/../ml_unigen_b200/csrc"
mkdir -p ../../build/exp
python build.py > /dev/null
for v in "$@"; do
  name=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include $v -c gemm_tc.cu -o ../../build/exp/gt_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a --shared -Xcompiler -fPIC -o ../../build/exp/lib_$name.so ug_capi.o decode_step.o decode_fast.o backbone_ops.o omni_attention.o magvit_ops.o ../../build/exp/gt_$name.o -lcuda
  echo built build/exp/lib_$name.so
done
