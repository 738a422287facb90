#!/bin/bash
# 2 GPUs: multi-GPU tests, the single-process device group (one plugin call over 2 PCIe links), torchrun bench N = 2 (both arms)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi2_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_shards.py -m gpu -x -q > gpurun_out/pytest_multi2.log 2>&1; echo "pytest multi rc $?"; tail -3 gpurun_out/pytest_multi2.log
timeout 900 python profiles/e2e_devices.py 1e8 32 > gpurun_out/e2e_devices_r02.jsonl 2> gpurun_out/e2e_devices_r02.err; cat gpurun_out/e2e_devices_r02.jsonl | cut -c1-300; tail -2 gpurun_out/e2e_devices_r02.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_C2_n2.json 2> gpurun_out/bench_C2_n2.err; echo "bench n2 rc $?"; python -c "
import json
d=json.load(open('gpurun_out/bench_C2_n2.json')); print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])"
