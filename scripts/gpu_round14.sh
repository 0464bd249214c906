#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_short.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_short.json')); print(d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['timed_region_ms'], d['e2e']['value'])"
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_short2.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_short2.json')); print(d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['timed_region_ms'], d['e2e']['value'])"
