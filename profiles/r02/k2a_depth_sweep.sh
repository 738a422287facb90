#!/bin/bash
mkdir -p gpurun_out
K2B_LIB=profiles/_ab/lib_k2a_sweep.so timeout 600 python profiles/k2a_depth_sweep.py 2>&1 | tee gpurun_out/k2a_depth_sweep.txt | grep BEST
