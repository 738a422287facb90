#!/bin/bash
# ncu --set full of the rolling pass-C kernel (C4 shape, 2e7 rows) and the grouped moments kernel (C3 shape, 2e7 rows); f64 K2a timings
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:online_main -s 1 -c 1 -o gpurun_out/k6_r02 -f python profiles/run_online.py 20000000 > gpurun_out/ncu_k6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:group_moments -s 1 -c 1 -o gpurun_out/k5_r02 -f python profiles/run_grouped.py 2e7 > gpurun_out/ncu_k5.log 2>&1
python profiles/k2a_bench.py > gpurun_out/k2a_r02.txt 2>&1; cat gpurun_out/k2a_r02.txt
ls -la gpurun_out/*.ncu-rep
