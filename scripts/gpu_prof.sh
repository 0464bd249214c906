#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_k1d.py --order sorted --columns 262144 --spl 4 --launches 3 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1d_tile -s 1 -c 1 -f -o gpurun_out/k1d_r1l python scripts/profile_k1d.py --order sorted --columns 262144 --spl 4 --launches 3 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/prof_plain.log; tail -1 gpurun_out/ncu_full.log
