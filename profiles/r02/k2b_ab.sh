#!/bin/bash
# K2b A/B: round-1 lane layout (kept build) without / with the pad-lane clamp vs the 16+16 lane layout (current build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_moments.py tests/test_gpu_frame.py -m gpu -x -q > gpurun_out/pytest_k2b.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_k2b.log
: > gpurun_out/k2b_ab.jsonl
run() { echo "== $*"; env "$@" python profiles/k2b_time.py 1e8 32 10 | tee -a gpurun_out/k2b_ab.jsonl; }
run K2B_LIB=profiles/_ab/lib_clamp.so PDSB_TC_CLAMP=0
run K2B_LIB=profiles/_ab/lib_clamp.so PDSB_TC_CLAMP=1
run K2B_NEW=1
run K2B_NEW=1 PDSB_TC_NCONV=3
run K2B_NEW=1 PDSB_TC_MODE=3
run K2B_NEW=1 PDSB_TC_MODE=3 PDSB_TC_NCONV=3
echo "== p=64"; python profiles/k2b_time.py 5e7 64 10 | tee -a gpurun_out/k2b_ab.jsonl
echo "== p=8";  python profiles/k2b_time.py 2e8 8 10 | tee -a gpurun_out/k2b_ab.jsonl
