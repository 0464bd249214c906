#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; cat gpurun_out/bench_r1.json
timeout 900 python bench.py --order hash --cpu-seconds 0 --steps 20 > gpurun_out/bench_r1_hash.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_r1_hash.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r1.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref_r1.json
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_short.json 2>> gpurun_out/bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
# first timed launch: 8 pre-roll launches + 3 warm-up launches precede it
ncu --set full --clock-control none --import-source on -k regex:k1d_tile -s 11 -c 1 -f -o gpurun_out/k1d_r1h python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
