#!/bin/bash
# timing probe of library variants (scripts/build_variant.sh): scripts/gpu_variants.sh name1 name2 ...   ("default" = the in-tree library)
mkdir -p gpurun_out
for v in "$@"; do
  lib=variants/libkid_$v.so; [ "$v" = default ] && lib=kinematicdriver.jl_b200/csrc/libkid_cuda.so
  echo "== $v" | tee -a gpurun_out/variants.log
  KID_CUDA_LIB=$PWD/$lib timeout 300 python scripts/profile_k1d.py --order morton --ic device --columns 262144 --spl 10 --launches 3 2>&1 | tail -1 | tee -a gpurun_out/variants.log
done
