#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_moments.py tests/test_gpu_frame.py -m gpu -x -q > gpurun_out/pytest_k2b.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_k2b.log
: > gpurun_out/k2b_v7.jsonl
for a in "1e8 32" "5e7 64" "5e7 48" "1e8 16" "2e8 8"; do timeout 300 python profiles/k2b_time.py $a 10 | tee -a gpurun_out/k2b_v7.jsonl | cut -c1-330; done
export K2B_LIB=profiles/_ab/lib_trace.so
PDSB_TC_DBG=16 timeout 300 python profiles/k2b_trace.py 5e7 32 > gpurun_out/k2b_trace_v7_p32.json 2> gpurun_out/k2b_trace_v7.err; python -c "
import json
d=json.load(open('gpurun_out/k2b_trace_v7_p32.json'))
for k,v in d.items(): print(k, v if not isinstance(v,dict) else v['mean'])"; tail -3 gpurun_out/k2b_trace_v7.err
for d in 0 8 15; do PDSB_TC_DBG=$d timeout 300 python profiles/k2b_time.py 5e7 32 10 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dbg', d['env'].get('PDSB_TC_DBG'), 'frame_ms', round(d['frame_ms'],3), 'col_ms', round(d['colmajor_ms'],3))"; done
unset K2B_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped" > gpurun_out/pytest_k5.log 2>&1; echo "pytest k5 rc $?"; tail -3 gpurun_out/pytest_k5.log
for st in 1 0; do PDSB_K5_STAGED=$st timeout 300 python bench.py --config C3 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_st$st.json 2>gpurun_out/bench_c3_st$st.err; python -c "import json;d=json.load(open('gpurun_out/bench_c3_st$st.json'));print('K5 staged=$st', d['ms_per_step'], d['roofline']['frac'], d['parity'])"; done
