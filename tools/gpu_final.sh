#!/bin/bash
# round-end validation: tests, smoke, default bench (all legs), reference arm, C2-size bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/$1
mkdir -p "$O"
( time python -m pytest tests -m gpu -q -x ) > "$O/pytest_gpu.log" 2>&1; tail -4 "$O/pytest_gpu.log"
python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
( time python bench.py ) > "$O/bench.json" 2> "$O/bench.err"; tail -4 "$O/bench.err"
( time python bench.py --impl reference --steps 2 --warmup 1 ) > "$O/bench_ref.json" 2> "$O/bench_ref.err"; tail -4 "$O/bench_ref.err"
python bench.py --steps 5 --warmup 3 --hii-dim 256 --box-len 300 --r-bubble-max 15 --no-cpu-baseline > "$O/bench_256.json" 2> "$O/bench_256.err"; tail -2 "$O/bench_256.err"
python - <<PY
import json
for f in ("bench", "bench_256"):
    d=json.load(open('$O/'+f+'.json'))
    print(f, 'value %.4g ms %.2f' % (d['value'], d['ms_per_step']), d['config']['workload'], 'e2e', d.get('e2e'))
    print('  ', d['roofline']); print('  ', d['step_roofline']); print('  ', d['cpu_baseline']); print('  ', d['clocks'], d['gpu_launches'])
    print('  ', {k: round(v,3) for k,v in d['kernel_profile_ms_per_step'].items()})
r=json.load(open('$O/bench_ref.json')); print('ref', r.get('value'), r.get('ms_per_step'), r.get('config'))
PY
