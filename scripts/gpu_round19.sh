#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_output.py > gpurun_out/bench_output.jsonl 2> gpurun_out/bench_output.err; echo rc=$?; cat gpurun_out/bench_output.jsonl; tail -3 gpurun_out/bench_output.err
timeout 600 python scripts/bench_output.py --columns 64 --steps 1000 > gpurun_out/bench_output_small.jsonl 2>> gpurun_out/bench_output.err; cat gpurun_out/bench_output_small.jsonl
timeout 900 python -m pytest tests/test_gpu_calibration.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
