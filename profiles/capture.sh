#!/bin/bash
# Round-1 profile capture (run under gpurun from the repo root):  bash profiles/capture.sh
# 1) launch list of one short bench run (device time of every kernel launch; cold-cache, serialised -> compare SHARES)
# 2) one `ncu --set full` capture of the dominant kernel (the tcgen05 Gram kernel) on a 2e7 x 32 frame
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_tcgen05 -s 2 -c 1 -o gpurun_out/prof_gram_r01 -f \
    python profiles/run_moments.py 2e7 32 4 > gpurun_out/prof_gram_r01.log 2>&1
