#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_moments.py tests/test_gpu_parity.py tests/test_golden_vectors.py -m gpu -x -q > gpurun_out/pytest_k2a.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_k2a.log
for k in 2 3; do echo "PDSB_K2A_KERNEL=$k"; PDSB_K2A_KERNEL=$k K2A_F64_ONLY=1 timeout 300 python profiles/k2a_bench.py; done 2>&1 | tee gpurun_out/k2a_f64_async.txt
