#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_calibration.py -m gpu -q -x > gpurun_out/pytest_cal.log 2>&1
tail -30 gpurun_out/pytest_cal.log | cut -c1-250
