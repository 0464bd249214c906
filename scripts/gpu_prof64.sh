#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_k1d.py --order sorted --columns 65536 --ft f64 --spl 4 --launches 3 > gpurun_out/prof_plain64.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1d_tile -s 1 -c 1 -f -o gpurun_out/k1d_f64 python scripts/profile_k1d.py --order sorted --columns 65536 --ft f64 --spl 4 --launches 3 > gpurun_out/ncu_full64.log 2>&1
cat gpurun_out/prof_plain64.log; tail -1 gpurun_out/ncu_full64.log
