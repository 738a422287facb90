#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_vectors.py tests/test_gpu_moments.py -m gpu -x -q > gpurun_out/pytest_k5.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_k5.log
for st in 0 1; do PDSB_K5_STAGED=$st timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_st$st.json 2>gpurun_out/bench_c3_st$st.err; python -c "import json;d=json.load(open('gpurun_out/bench_c3_st$st.json'));print('K5 staged=$st', d['ms_per_step'], d['roofline']['frac'], d['parity'])"; done
for a in "2e8 8" "1e8 10"; do timeout 300 python profiles/k2b_time.py $a 10 | cut -c1-260; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_C3.csv python bench.py --config C3 --steps 3 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; grep -E "group_moments|group_solve|item_rows" gpurun_out/launches_C3.csv | tail -4 | cut -c1-60,200-
