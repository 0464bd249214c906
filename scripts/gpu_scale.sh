#!/bin/bash
# usage: gpu_scale.sh N   (run under gpurun --gpus N)
mkdir -p gpurun_out
n=$1
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 40 --warmup 4 --cpu-seconds 0 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "n=$n rc=$?"; tail -2 gpurun_out/bench_n$n.err | cut -c1-200; cat gpurun_out/bench_n$n.json
