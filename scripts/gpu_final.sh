#!/bin/bash
# Round-end evidence: tests, smoke, bench lines (sorted / hash / reference arm / 600-step windows), launch list, ncu capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -2 gpurun_out/bench_n1.err
timeout 900 python bench.py --order hash --cpu-seconds 0 > gpurun_out/bench_n1_hash.json 2> gpurun_out/bench_n1_hash.err; echo rc=$?
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
timeout 900 python bench.py --steps 600 --t-start 0 --cpu-seconds 0 > gpurun_out/bench_n1_w0_600.json 2> gpurun_out/bench_w0.err; echo rc=$?
timeout 900 python bench.py --steps 600 --t-start 1800 --cpu-seconds 0 > gpurun_out/bench_n1_w1800_2400.json 2> gpurun_out/bench_w1800.err; echo rc=$?
for f in bench_n1 bench_n1_hash bench_ref bench_n1_w0_600 bench_n1_w1800_2400; do python -c "
import json;d=json.load(open('gpurun_out/$f.json'));print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('clocks'))"; done
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_short.json 2>> gpurun_out/bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-300
timeout 900 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; echo rc=$?; tail -2 gpurun_out/configs.err; head -c 1500 gpurun_out/configs.json
