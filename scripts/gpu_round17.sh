#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 900 python bench.py --order hash --cpu-seconds 0 > gpurun_out/bench_n1_hash.json 2> gpurun_out/bench_n1_hash.err; echo rc=$?; cat gpurun_out/bench_n1_hash.json
timeout 900 python bench.py --steps 600 --t-start 0 --cpu-seconds 0 > gpurun_out/bench_n1_w0_600.json 2> gpurun_out/bench_w0.err; echo rc=$?; cat gpurun_out/bench_n1_w0_600.json
timeout 900 python bench.py --steps 600 --t-start 1800 --cpu-seconds 0 > gpurun_out/bench_n1_w1800_2400.json 2> gpurun_out/bench_w1800.err; echo rc=$?; cat gpurun_out/bench_n1_w1800_2400.json
