#!/bin/bash
# grouped register kernel: precomputed item rows (bit 0) and a work queue over a persistent grid (bit 1), C3
mkdir -p gpurun_out
for m in 0 1 2 3; do for ch in 8192 4096; do PDSB_K5_MODE=$m PDSB_K5_CHUNK=$ch timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_m$m.json 2>gpurun_out/bench_c3_m$m.err; python -c "import json;d=json.load(open('gpurun_out/bench_c3_m$m.json'));print('K5 mode=$m chunk=$ch', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity']['coef_rel_err_vs_numpy_lstsq_f64_50_random_groups'])"; done; done 2>&1 | tee gpurun_out/k5_mode_sweep.txt
PDSB_K5_MODE=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_vectors.py -m gpu -x -q 2>&1 | tail -2
