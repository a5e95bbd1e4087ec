#!/bin/bash
# Whole GPU suite + smoke + one bench line.   usage: gpurun --timeout 1800 -- 'bash tools/gpu_suite.sh <tag> [pytest -k expr]'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-suite}
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi -L > "$O/gpu.txt"
if [ -n "$2" ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=10 -k "$2" ) > "$O/pytest_gpu.log" 2>&1
else
  ( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=10 ) > "$O/pytest_gpu.log" 2>&1
fi
tail -22 "$O/pytest_gpu.log"
python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
python bench.py --steps 10 --warmup 3 --no-yardstick > "$O/bench.json" 2> "$O/bench.err"; tail -2 "$O/bench.err" | cut -c1-300; cut -c1-1200 "$O/bench.json"
