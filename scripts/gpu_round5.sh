#!/bin/bash
mkdir -p gpurun_out
for v in 512 1024 1024_asqrt; do
  echo "== variant $v"; KID_CUDA_LIB=$PWD/build_variants/lib_$v.so timeout 300 python scripts/profile_k1d.py --launches 4
done > gpurun_out/variants.log 2>&1
echo "== 1024 nw=16" >> gpurun_out/variants.log; KID_CUDA_LIB=$PWD/build_variants/lib_1024.so timeout 300 python scripts/profile_k1d.py --nw 16 >> gpurun_out/variants.log 2>&1
echo "== 1024 t=1800" >> gpurun_out/variants.log; KID_CUDA_LIB=$PWD/build_variants/lib_1024.so timeout 300 python scripts/profile_k1d.py --t-start 1800 >> gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
KID_CUDA_LIB=$PWD/build_variants/lib_1024_asqrt.so timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_asqrt.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_asqrt.log | tail; grep -E "AssertionError: \(" gpurun_out/pytest_gpu_asqrt.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail; grep -E "AssertionError: \(" gpurun_out/pytest_gpu.log | cut -c1-300
