#!/bin/bash
# tests + smoke + full default bench (device, e2e, cpu baseline) + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/$1
mkdir -p "$O"
( time python -m pytest tests -m gpu -q -x ) > "$O/pytest_gpu.log" 2>&1; tail -4 "$O/pytest_gpu.log"
python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
( time python bench.py ) > "$O/bench.json" 2> "$O/bench.err"; tail -4 "$O/bench.err"
( time python bench.py --impl reference --steps 2 --warmup 1 ) > "$O/bench_ref.json" 2> "$O/bench_ref.err"; tail -4 "$O/bench_ref.err"
python - <<PY
import json
d=json.load(open('$O/bench.json'))
print('value %.4g ms %.2f' % (d['value'], d['ms_per_step']), 'e2e', d.get('e2e'))
print(d['roofline']); print(d['step_roofline']); print(d['cpu_baseline']); print(d['clocks'])
print({k: round(v,3) for k,v in d['kernel_profile_ms_per_step'].items()})
r=json.load(open('$O/bench_ref.json')); print('ref', r.get('value'), r.get('ms_per_step'), r.get('cpu_baseline',{}).get('sample'))
PY
