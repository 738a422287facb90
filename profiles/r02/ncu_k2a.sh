#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gram_dmma -s 1 -c 1 -o gpurun_out/k2a_side_p32 -f python profiles/k2a_one.py 2e7 32 > gpurun_out/ncu_k2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_dmma -s 1 -c 1 -o gpurun_out/k2a_wide_p30 -f python profiles/k2a_one.py 2e7 30 >> gpurun_out/ncu_k2a.log 2>&1
ls -la gpurun_out/*.ncu-rep
