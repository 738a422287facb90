python -m pytest tests/test_gpu_parity.py tests/test_golden_vectors.py tests/test_gpu_moments.py -m gpu -q > gpurun_out/pytest_gpu_r02f.log 2>&1; tail -6 gpurun_out/pytest_gpu_r02f.log
for v in 0 1; do PDSB_K6_VAR=$v python bench.py --config C4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c4_var$v.json 2>gpurun_out/bench_c4_var$v.err; python -c "import json;d=json.load(open('gpurun_out/bench_c4_var$v.json'));print('K6 var',$v, d['ms_per_step'], d['roofline']['frac'], d['parity'])"; done
for k in 16 8; do PDSB_K2A_KERNEL=$k python profiles/k2a_bench.py > gpurun_out/k2a_kern$k.txt 2>&1; echo kernel $k; cat gpurun_out/k2a_kern$k.txt; done
python bench.py --config C3 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3.json 2>gpurun_out/bench_c3.err; python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print('K5', d['ms_per_step'], d['roofline']['frac'], d['parity'])"
bash profiles/sanitize.sh
