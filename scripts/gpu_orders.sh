#!/bin/bash
mkdir -p gpurun_out
for o in hash sorted uniform; do timeout 300 python scripts/profile_k1d.py --launches 4 --order $o; done > gpurun_out/orders.log 2>&1
for o in hash sorted; do timeout 300 python scripts/profile_k1d.py --launches 4 --order $o --t-start 1800; done >> gpurun_out/orders.log 2>&1
cat gpurun_out/orders.log
