#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "n=$n rc=$?"; tail -2 gpurun_out/bench_n$n.err | cut -c1-200; cat gpurun_out/bench_n$n.json
done
