#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/k2d_scale.jsonl
python scripts/bench_k2d.py 2>/dev/null | tail -1 >> gpurun_out/k2d_scale.jsonl
for n in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n scripts/bench_k2d.py 2>/dev/null | tail -1 >> gpurun_out/k2d_scale.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n scripts/bench_k2d.py --weak 2>/dev/null | tail -1 >> gpurun_out/k2d_scale.jsonl
done
cat gpurun_out/k2d_scale.jsonl
