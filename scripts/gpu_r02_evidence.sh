#!/bin/bash
# Round-2 evidence: GPU tests, smoke, bench lines (ours / reference arm / hash order), launch list, ncu --set full of the step kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -2 gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference --steps 40 --warmup 4 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
timeout 900 python bench.py --order hash --cpu-seconds 0 --k2d-steps 0 > gpurun_out/bench_n1_hash.json 2> gpurun_out/bench_n1_hash.err; echo rc=$?
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-parity --k2d-steps 0 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-parity --k2d-steps 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1d_pair -s 9 -c 1 -f -o gpurun_out/k1d_pair_r02 python bench.py --steps 20 --warmup 3 --cpu-seconds 0 --no-parity --k2d-steps 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
mkdir -p gpurun_out/sum
python scripts/ncu_summary.py gpurun_out/k1d_pair_r02.ncu-rep > gpurun_out/sum/k1d_pair_f32_2M.json
python scripts/ncu_sass_hist.py gpurun_out/k1d_pair_r02.ncu-rep 40 > gpurun_out/sum/k1d_pair_f32_2M_sass.txt
ncu -i gpurun_out/k1d_pair_r02.ncu-rep --page details > gpurun_out/sum/k1d_pair_f32_2M_details.txt 2>/dev/null
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log | cut -c1-300
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/f32x2_peak scripts/microbench/f32x2_peak.cu 2>/dev/null && /tmp/f32x2_peak > gpurun_out/f32x2_peak.jsonl 2>&1; tail -3 gpurun_out/f32x2_peak.jsonl
for f in bench_n1 bench_n1_hash bench_ref; do python -c "
import json;d=json.load(open('gpurun_out/$f.json'));print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('clocks'), d.get('parity_sample'), (d.get('secondary') or {}).get('k2d'))"; done
ls -la gpurun_out
