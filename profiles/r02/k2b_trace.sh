#!/bin/bash
mkdir -p gpurun_out
export K2B_LIB=profiles/_ab/lib_trace.so
for p in 32 8; do PDSB_TC_DBG=16 python profiles/k2b_trace.py 5e7 $p > gpurun_out/k2b_trace_p$p.json 2> gpurun_out/k2b_trace_p$p.err; cat gpurun_out/k2b_trace_p$p.json; tail -3 gpurun_out/k2b_trace_p$p.err; done
: > gpurun_out/k2b_abl.jsonl
for d in 0 1 2 4 8 7 15; do PDSB_TC_DBG=$d python profiles/k2b_time.py 5e7 32 10 | tee -a gpurun_out/k2b_abl.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dbg', d['env'].get('PDSB_TC_DBG'), 'frame_ms', round(d['frame_ms'],3), 'col_ms', round(d['colmajor_ms'],3))"; done
unset K2B_LIB
python profiles/k2b_time.py 1e8 32 10 | tee gpurun_out/k2b_new2.json
