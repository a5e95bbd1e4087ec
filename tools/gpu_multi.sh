#!/bin/bash
# usage: gpurun --gpus N -- bash tools/gpu_multi.sh <tag> <N>
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/$1; N=${2:-2}
mkdir -p "$O"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
for mode in boxes radius; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --partition $mode > "$O/bench_${mode}_$N.json" 2> "$O/bench_${mode}_$N.err"
  tail -3 "$O/bench_${mode}_$N.err"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_${mode}_$N.json') if l.startswith('{')][-1])
    print('$mode', 'N=', d['n_gpus'], '%.4g cells/s' % d['value'], 'ms %.2f' % d['ms_per_step'], d['scaling'], d['config']['ms_perturb'], d['config']['ms_ionize'], d['config']['global_xH'])
except Exception as e:
    print('$mode ERR', e)
PY
done
