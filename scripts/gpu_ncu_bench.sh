#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_short.json 2> gpurun_out/bench.err && \
ncu --set full --clock-control none --import-source on -k regex:k1d_tile -s 11 -c 1 -f -o gpurun_out/k1d_bench_r1 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
