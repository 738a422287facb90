#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu.log
for c in C2 C4 C5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${c}_b.json 2> gpurun_out/bench_${c}_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_b.json')); print('$c', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'kms', round(d['roofline']['kernel_ms'],3), d['parity'])"; done
for a in "2e8 8" "1e8 4" "1e8 10"; do timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-260; done
for k in 16 8 0; do echo "PDSB_K2A_KERNEL=$k"; PDSB_K2A_KERNEL=$k timeout 300 python profiles/k2a_bench.py 2>&1 | tail -4; done
bash profiles/sanitize.sh
