#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail; grep -E "AssertionError: \(|Error" gpurun_out/pytest_gpu.log | cut -c1-300 | head
for o in hash sorted; do timeout 300 python scripts/profile_k1d.py --launches 4 --order $o; done > gpurun_out/orders.log 2>&1
timeout 300 python scripts/profile_k1d.py --launches 4 --order sorted --ft f64 --columns 16384 >> gpurun_out/orders.log 2>&1
cat gpurun_out/orders.log
