#!/bin/bash
mkdir -p gpurun_out
P="timeout 300 python scripts/profile_k1d.py --order sorted --columns 262144 --spl 4 --launches 3"
{ $P; $P --no-fixed 1; $P --no-fixed 1 --tile-c 16 --nw 32; $P --no-fixed 1 --tile-c 16 --nw 64; $P --no-fixed 1 --tile-c 8 --nw 32; } > gpurun_out/prof_tiles.log 2>&1
cat gpurun_out/prof_tiles.log
