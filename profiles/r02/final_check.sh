#!/bin/bash
# round-end state: full GPU suite, smoke, one bench line per BASELINE config (C5 device-resident only), the CPU arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke_final.log
rm -f gpurun_out/bench_state_3.jsonl
for c in C2 C1 C3 C4; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc $?"; cat gpurun_out/bench_$c.json >> gpurun_out/bench_state_3.jsonl
done
timeout 600 python bench.py --config C5 --no-e2e --no-cpu > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err; echo "bench C5 rc $?"; cat gpurun_out/bench_C5.json >> gpurun_out/bench_state_3.jsonl
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc $?"; cat gpurun_out/bench_reference.json >> gpurun_out/bench_state_3.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/bench_state_3.jsonl'):
    try:
        d = json.loads(l)
    except Exception as e:
        print('bad line', e); continue
    r = d.get('roofline') or {}
    e2e = d.get('e2e') or {}
    print(d.get('impl', 'ours'), d['config'].get('workload', '')[:40], '| ms', round(d.get('ms_per_step', 0), 3), '| value %.3g' % d['value'], '| frac', r.get('frac') and round(r['frac'], 3),
          '| e2e', e2e.get('value') and '%.3g' % e2e['value'], '| cpu', (d.get('cpu_baseline') or {}).get('value'), '| clocks', d.get('clocks', {}).get('sm_mhz'), '| parity', d.get('parity'))
PY
