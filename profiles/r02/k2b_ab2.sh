#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_moments.py tests/test_gpu_frame.py -m gpu -x -q > gpurun_out/pytest_k2b.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_k2b.log
for rep in 1 2; do
for a in "1e8 32" "5e7 64"; do
  echo "new:"; timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-200
  echo "old:"; K2B_LIB=profiles/_ab/lib_trace.so timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-200
done; done
