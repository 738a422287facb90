#!/bin/bash
mkdir -p gpurun_out
for k in 2; do echo "PDSB_K2A_KERNEL=$k"; PDSB_K2A_KERNEL=$k K2A_F64_ONLY=1 timeout 300 python profiles/k2a_bench.py; done 2>&1 | tee gpurun_out/k2a_f64_wide2.txt
