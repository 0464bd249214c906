#!/bin/bash
# ncu --set full of one launch (10 fused steps, 262,144 columns, developed state) of the step kernel: scripts/gpu_ncu_pair.sh NAME [lib]
mkdir -p gpurun_out
name=$1; lib=${2:-kinematicdriver.jl_b200/csrc/libkid_cuda.so}
export KID_CUDA_LIB=$PWD/$lib
timeout 300 python scripts/profile_k1d.py --order morton --ic device --columns 262144 --spl 10 --launches 3 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1d_pair -s 1 -c 1 -f -o gpurun_out/$name python scripts/profile_k1d.py --order morton --ic device --columns 262144 --spl 10 --launches 3 > gpurun_out/ncu_$name.log 2>&1
tail -2 gpurun_out/ncu_$name.log | cut -c1-200
