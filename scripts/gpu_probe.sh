#!/bin/bash
# quick turn-around check after a kernel change: three timing probes, then the whole GPU test suite
mkdir -p gpurun_out
timeout 300 python scripts/profile_k1d.py --order sorted --columns 262144 --spl 4 --launches 3 > gpurun_out/prof_plain.log 2>&1
timeout 300 python scripts/profile_k1d.py --order hash --columns 65536 --spl 4 --launches 3 >> gpurun_out/prof_plain.log 2>&1
timeout 300 python scripts/profile_k1d.py --order sorted --columns 65536 --ft f64 --spl 4 --launches 3 >> gpurun_out/prof_plain.log 2>&1
cat gpurun_out/prof_plain.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail; grep -E "AssertionError: \(|Error" gpurun_out/pytest_gpu.log | cut -c1-300 | head
