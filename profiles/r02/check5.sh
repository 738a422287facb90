#!/bin/bash
# grouped path: items of equal size per group, chunk sweep (C3), small-p whole-frame kernel unchanged?
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_vectors.py tests/test_gpu_moments.py tests/test_gpu_frame.py -m gpu -x -q > gpurun_out/pytest_k5.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_k5.log
for ch in 8192 4096 2048; do PDSB_K5_CHUNK=$ch timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_ch$ch.json 2>gpurun_out/bench_c3_ch$ch.err; python -c "import json;d=json.load(open('gpurun_out/bench_c3_ch$ch.json'));print('K5 chunk=$ch', d['ms_per_step'], d['roofline']['frac'], d['parity'])"; done
for a in "2e8 8" "1e8 10"; do timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-260; done
