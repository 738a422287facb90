#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_moments.py tests/test_gpu_frame.py -m gpu -x -q > gpurun_out/pytest_k2b.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_k2b.log
for rep in 1 2; do
for a in "1e8 32" "5e7 64"; do
  echo "convpipe+2sum:"; timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-200
  echo "sidewarps+2sum:"; K2B_LIB=profiles/_ab/lib_sidew.so timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-200
  echo "head (f64 epilogue):"; K2B_LIB=profiles/_ab/lib_head.so timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-200
done; done
export K2B_LIB=profiles/_ab/lib_trace.so
PDSB_TC_DBG=16 timeout 300 python profiles/k2b_trace.py 5e7 32 > gpurun_out/k2b_trace_v7_p32.json 2> gpurun_out/k2b_trace_v7.err; python -c "
import json
d=json.load(open('gpurun_out/k2b_trace_v7_p32.json'))
for k,v in d.items(): print(k, v if not isinstance(v,dict) else v['mean'])"; tail -3 gpurun_out/k2b_trace_v7.err
