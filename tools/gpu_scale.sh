#!/bin/bash
# The driver's scaling launch at N ranks (default partition = boxes + strong record), plus the slab leg with host timing.
# usage: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_scale.sh <tag> <N>'
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/$1; N=${2:-2}
mkdir -p "$O"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > "$O/gpu.txt"
nvidia-smi topo -m > "$O/topo.txt" 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 5 --warmup 3 ) > "$O/bench_$N.json" 2> "$O/bench_$N.err"; tail -4 "$O/bench_$N.err" | cut -c1-400
python - "$O/bench_$N.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('N=', d['n_gpus'], '%.4g cells/s' % d['value'], 'ms %.2f' % d['ms_per_step'], d['scaling'], 'e2e', d.get('e2e', {}).get('ms_per_step'))
    for k, v in (d.get('strong') or {}).items():
        print(' strong', k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items()})
except Exception as e:
    print('ERR', e)
PY
B200_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus $N --steps 5 --warmup 3 --partition slab --no-e2e > "$O/bench_slab_$N.json" 2> "$O/bench_slab_$N.err"
python - "$O/bench_slab_$N.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('slab N=', d['n_gpus'], 'ms %.2f' % d['ms_per_step'], d['config']['ms_perturb'], d['config']['ms_ionize'], {k: round(v, 2) for k, v in d['kernel_profile_ms_per_step'].items()})
except Exception as e:
    print('ERR', e)
PY
grep -m3 "ionize host" "$O/bench_slab_$N.err"
