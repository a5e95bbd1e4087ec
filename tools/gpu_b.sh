#!/bin/bash
# parity tests + default bench (device leg only) with per-kernel ms; usage: gpu_b.sh <tag> [ENV=val ...]
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=$1; shift
O=gpurun_out/$TAG
mkdir -p "$O"
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > "$O/pytest_gpu.log"; tail -3 "$O/pytest_gpu.log"
[ $# -eq 0 ] && set -- B200_NOP=0
i=0
for v in "$@"; do
  i=$((i+1))
  env $v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > "$O/v$i.json" 2> "$O/v$i.err"
  tail -2 "$O/v$i.err"
  python - <<PY
import json
try:
    d=json.load(open('$O/v$i.json'))
    print('[$v]', '%.4g cells/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'perturb %.2f ionize %.2f' % (d['config']['ms_perturb'], d['config']['ms_ionize']), 'xH %.7f' % d['config']['global_xH'])
    print('  ', {k: round(v,3) for k,v in d['kernel_profile_ms_per_step'].items()})
except Exception as e:
    print('[$v] ERR', e)
PY
done
