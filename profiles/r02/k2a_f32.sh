#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_moments.py tests/test_gpu_parity.py tests/test_golden_vectors.py tests/test_gpu_models.py -m gpu -x -q > gpurun_out/pytest_k2a.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_k2a.log
K2A_F64_ONLY=1 timeout 300 python profiles/k2a_bench.py 2>&1 | tee gpurun_out/k2a_f32_dmma.txt
