#!/bin/bash
# (reports are summarised on the box: gpurun brings back at most 64 MiB)
# ncu --set full of the other kernel families: Float64 tile kernel (NonEq+2M, NonEq+1M), Box step kernel (f64, f32), K2D kernels
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 python scripts/profile_k1d.py --order sorted --columns 65536 --ft f64 --spl 4 --launches 3 | tail -1
$N -k regex:k1d_tile -s 1 -c 1 -o gpurun_out/tile_f64_2M python scripts/profile_k1d.py --order sorted --columns 65536 --ft f64 --spl 4 --launches 3 > gpurun_out/ncu_f.log 2>&1
timeout 300 python scripts/profile_k1d.py --order sorted --columns 65536 --ft f64 --precip Precipitation1M --spl 4 --launches 3 | tail -1
$N -k regex:k1d_tile -s 1 -c 1 -o gpurun_out/tile_f64_1M python scripts/profile_k1d.py --order sorted --columns 65536 --ft f64 --precip Precipitation1M --spl 4 --launches 3 >> gpurun_out/ncu_f.log 2>&1
timeout 300 python scripts/profile_k1d.py --order sorted --columns 65536 --ft f32 --precip Precipitation1M --spl 4 --launches 3 | tail -1
$N -k regex:k1d_tile -s 1 -c 1 -o gpurun_out/tile_f32_1M python scripts/profile_k1d.py --order sorted --columns 65536 --ft f32 --precip Precipitation1M --spl 4 --launches 3 >> gpurun_out/ncu_f.log 2>&1
timeout 300 python scripts/profile_box.py --ft f64 | tail -1
$N -k regex:box_step -s 1 -c 1 -o gpurun_out/box_f64_1M python scripts/profile_box.py --ft f64 >> gpurun_out/ncu_f.log 2>&1
timeout 300 python scripts/profile_box.py --ft f32 | tail -1
$N -k regex:box_step -s 1 -c 1 -o gpurun_out/box_f32_1M python scripts/profile_box.py --ft f32 >> gpurun_out/ncu_f.log 2>&1
$N -k regex:k2d_hdiv -s 100 -c 1 -o gpurun_out/k2d_hdiv python scripts/bench_k2d.py --steps 30 >> gpurun_out/ncu_f.log 2>&1
$N -k regex:k2d_dss -s 100 -c 1 -o gpurun_out/k2d_dss python scripts/bench_k2d.py --steps 30 >> gpurun_out/ncu_f.log 2>&1
$N -k regex:k1d_tile -s 100 -c 1 -o gpurun_out/k2d_tile python scripts/bench_k2d.py --steps 30 >> gpurun_out/ncu_f.log 2>&1
mkdir -p gpurun_out/sum
for r in gpurun_out/*.ncu-rep; do b=$(basename $r .ncu-rep); python scripts/ncu_summary.py $r > gpurun_out/sum/$b.json; python scripts/ncu_sass_hist.py $r 30 > gpurun_out/sum/${b}_sass.txt; rm -f $r; done
grep -c "Report" gpurun_out/ncu_f.log; ls -la gpurun_out/sum
