#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  echo "== variant $v"; KID_CUDA_LIB=$PWD/build_variants/lib_$v.so timeout 300 python scripts/profile_k1d.py --launches 4
  KID_CUDA_LIB=$PWD/build_variants/lib_$v.so timeout 300 python scripts/profile_k1d.py --launches 3 --spl 8
done > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
