#!/bin/bash
# GPU parity suite (incl. the recombination chain), smoke, and one device-leg bench line of the default workload
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-rec}
O=gpurun_out/$TAG
mkdir -p "$O"
( time timeout 240 python -m pytest tests -m gpu -q -x ) > "$O/pytest_gpu.log" 2>&1; tail -6 "$O/pytest_gpu.log"
timeout 60 python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > "$O/bench.json" 2> "$O/bench.err"; tail -2 "$O/bench.err"
python -c "
import json; d=json.load(open('$O/bench.json')); print('bench %.4g cells/s, %.2f ms/step' % (d['value'], d['ms_per_step']))"
