#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/k2d_multirank.py > gpurun_out/k2d_multirank.log 2>&1; echo rc=$?; grep -E "rank|Error|error" gpurun_out/k2d_multirank.log | tail
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
