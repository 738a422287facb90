#!/bin/bash
# Round-1 profile capture (run under gpurun from the repo root):  bash profiles/capture.sh
# 1) launch list of one short bench run (device time of every kernel launch; cold-cache, serialised -> compare SHARES)
# 2) one `ncu --set full` capture of the dominant kernel (the tcgen05 Gram kernel) at the BENCH size (1e8 x 33 frame):
#    duration, dram bytes (roofline.traffic), pipe utilisation
# 3) launch lists of the other configs' kernels (grouped, rolling, recursive)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_tcgen05 -s 3 -c 1 -o gpurun_out/prof_gram_r01 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_gram_r01.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches_online_r01.csv \
    python profiles/run_online.py 1e8 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 4 --csv --log-file gpurun_out/launches_grouped_r01.csv \
    python profiles/run_grouped.py 1e8 > /dev/null 2>&1
