#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do PDSB_K6_VAR=$v timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c4_var$v.json 2>gpurun_out/bench_c4_var$v.err; python -c "import json;d=json.load(open('gpurun_out/bench_c4_var$v.json'));print('K6 var',$v, d['ms_per_step'], d['roofline']['frac'], d['parity'])"; done
compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 3 python profiles/sanitize_target.py > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_r02.log; tail -3 gpurun_out/sanitizer_racecheck_r02.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rolling or recursive" > gpurun_out/pytest_k6.log 2>&1; echo "pytest k6 rc $?"; tail -2 gpurun_out/pytest_k6.log
