#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail; grep -E "AssertionError: \(" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_r1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r1.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref_r1.json
# launch list of the bench command (short: 8 steps), after the same command exited 0 without ncu
timeout 600 python bench.py --steps 8 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_short.json 2>> gpurun_out/bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 8 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
# full capture of the dominant kernel in the timed region (skip the pre-roll launch and 3 warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:k1d_tile -s 5 -c 1 -f -o gpurun_out/k1d_r1c python bench.py --steps 8 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
