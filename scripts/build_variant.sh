#!/bin/bash
# Experiment builds of libkid_cuda.so: scripts/build_variant.sh NAME "-DMACRO ..." -> variants/libkid_NAME.so
# (the Precipitation2M translation unit and the C ABI are recompiled with the extra flags; run with KID_CUDA_LIB=variants/libkid_NAME.so)
set -e
cd "$(dirname "$0")/../kinematicdriver.jl_b200/csrc"
name=$1; shift
mkdir -p /tmp/kidvar_$name ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -use_fast_math -prec-sqrt=true -std=c++17 -lineinfo --expt-relaxed-constexpr \
  -Xcompiler -fPIC -ccbin /usr/bin/g++ -diag-suppress 177 -Xptxas -v $@ -DKID_PRECIP=3 -c kid_inst.cu -o /tmp/kidvar_$name/kid_inst3.o 2> /tmp/kidvar_$name/ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -use_fast_math -prec-sqrt=true -std=c++17 -lineinfo --expt-relaxed-constexpr \
  -Xcompiler -fPIC -ccbin /usr/bin/g++ -diag-suppress 177 $@ -c kid_capi.cu -o /tmp/kidvar_$name/kid_capi.o
grep -A2 "k1d_pair_kernelILi1" /tmp/kidvar_$name/ptxas.log | grep -E "registers|spill" || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../variants/libkid_$name.so kid_inst0.o kid_inst1.o kid_inst2.o /tmp/kidvar_$name/kid_inst3.o /tmp/kidvar_$name/kid_capi.o
echo built variants/libkid_$name.so
