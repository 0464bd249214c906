#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail; grep -E "AssertionError: \(|Error" gpurun_out/pytest_gpu.log | cut -c1-300 | head
timeout 300 python scripts/profile_k1d.py --launches 4 > gpurun_out/prof_plain.log 2>&1
timeout 300 python scripts/profile_k1d.py --launches 4 --t-start 1800 >> gpurun_out/prof_plain.log 2>&1
timeout 300 python scripts/profile_k1d.py --launches 4 --ft f64 --columns 16384 >> gpurun_out/prof_plain.log 2>&1
cat gpurun_out/prof_plain.log
timeout 300 python scripts/profile_k1d.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1d_tile -s 1 -c 1 -f -o gpurun_out/k1d_r1d python scripts/profile_k1d.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
