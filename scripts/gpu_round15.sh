#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pipelined or state_roundtrip" > gpurun_out/pytest_pipe.log 2>&1
tail -5 gpurun_out/pytest_pipe.log
for ch in 1 4 8; do
timeout 600 python bench.py --steps 40 --warmup 4 --cpu-seconds 0 --e2e-chunks $ch > gpurun_out/bench_e2e_$ch.json 2> gpurun_out/bench_e2e_$ch.err; echo rc=$?; tail -3 gpurun_out/bench_e2e_$ch.err
python -c "
import json;d=json.load(open('gpurun_out/bench_e2e_$ch.json'));print($ch, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
