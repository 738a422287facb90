#!/bin/bash
mkdir -p gpurun_out
timeout 70 ncu --set full --clock-control none --import-source on -k regex:gram_dmma_side -s 1 -c 1 -o gpurun_out/k2a_side_p32_final -f python profiles/k2a_one.py 2e7 32 > gpurun_out/ncu_k2a_final.log 2>&1; echo "ncu rc $?"; ls -la gpurun_out/k2a_side_p32_final.ncu-rep
