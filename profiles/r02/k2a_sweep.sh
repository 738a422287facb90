#!/bin/bash
# side kernel at p = 32 (NB = 4): ring depth x CTAs per SM  (0 = shipped 3 x 3, 6 = 4 x 2, 1 = 5 x 2, 2 = 6 x 2, 3 = 7 x 2, 4 = 8 x 2, 5 = 8 x 1)
mkdir -p gpurun_out
for cfg in 0 6 1 2 3 4 5; do K2B_LIB=profiles/_ab/lib_k2a_sweep.so PDSB_K2A_CFG=$cfg timeout 120 python profiles/k2a_one.py 2e7 32; done 2>&1 | tee gpurun_out/k2a_sweep2.txt
