#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_k2d.py -m gpu -q > gpurun_out/pytest_k2d.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_k2d.log | tail -30; grep -E "Error|assert " gpurun_out/pytest_k2d.log | cut -c1-300 | head -30
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail; grep -E "AssertionError: \(" gpurun_out/pytest_gpu.log | cut -c1-300
