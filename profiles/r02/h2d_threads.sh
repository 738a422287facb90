#!/bin/bash
# pageable e2e of C2 against the number of staging threads (same box); CPU quota of the box
mkdir -p gpurun_out
echo "nproc $(nproc)"; echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; grep -E "nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo
for w in 6 10 12 4; do PDS_B200_H2D_THREADS=$w timeout 400 python bench.py --config C2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2_w$w.json 2> gpurun_out/bench_c2_w$w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_w$w.json')); e=d['e2e']; print('threads $w: pageable', round(e['ms_per_step'],1), 'ms, pinned', round(e['pinned_ms_per_step'],1), 'ms')"; grep -E "nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo; done 2>&1 | tee gpurun_out/h2d_threads2.txt
