#!/bin/bash
# quick GPU iteration: parity tests + device-leg bench at 256^3 and 512^3 (DIM=1024), per-kernel ms
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-q}
O=gpurun_out/$TAG
mkdir -p "$O"
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > "$O/pytest_gpu.log"; tail -5 "$O/pytest_gpu.log"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > "$O/bench_256.json" 2> "$O/bench_256.err"; tail -3 "$O/bench_256.err"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --hii-dim 512 --dim 1024 --box-len 768 --r-bubble-max 40 > "$O/bench_512.json" 2> "$O/bench_512.err"; tail -3 "$O/bench_512.err"
python - <<PY
import json
for f in ('bench_256','bench_512'):
    try:
        d=json.load(open('$O/'+f+'.json'))
        print(f, '%.4g cells/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'perturb %.2f ionize %.2f' % (d['config']['ms_perturb'], d['config']['ms_ionize']), 'xH', d['config']['global_xH'])
        print('  ', {k: round(v,3) for k,v in d['kernel_profile_ms_per_step'].items()})
        print('  ', d['roofline'])
    except Exception as e:
        print(f, 'ERR', e)
PY
