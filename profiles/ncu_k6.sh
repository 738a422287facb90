#!/bin/bash
# ncu of the rolling path (C4 shape at 2e7 rows): launch list + full capture of the pass-C kernel
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_online_r02.csv python profiles/run_online.py 20000000 > gpurun_out/run_online.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:online_main -s 1 -c 1 -o gpurun_out/k6_r02 -f python profiles/run_online.py 20000000 >> gpurun_out/run_online.log 2>&1
