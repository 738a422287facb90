#!/bin/bash
# Round-2 profile capture (run under gpurun from the repo root):  bash profiles/r02/capture.sh
# 1) launch lists (device time of every kernel launch; cold-cache, serialised -> compare SHARES) of one short bench run per config
# 2) one `ncu --set full` capture of the dominant kernel of C2 (the tcgen05 Gram kernel) at the BENCH size -> duration, dram bytes
#    (roofline.traffic), pipe utilisation; the same for C5's shape
mkdir -p gpurun_out
for c in C2 C3 C4 C5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$c.csv \
      python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:gram_tcgen05_kernel -s 3 -c 1 -o gpurun_out/prof_gram_r02 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_gram_r02.log 2>&1
ncu --set full --clock-control none -k regex:gram_tcgen05_kernel -s 3 -c 1 -o gpurun_out/prof_gram_c5_r02 -f \
    python bench.py --config C5 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_gram_c5_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
