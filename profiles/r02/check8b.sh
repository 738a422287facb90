#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_vectors.py tests/test_gpu_models.py tests/test_gpu_shards.py -m gpu -x -q > gpurun_out/pytest_k6.log 2>&1; grep -E "^E |Error|FAILED|passed|failed" gpurun_out/pytest_k6.log | head -30
