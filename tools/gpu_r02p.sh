#!/bin/bash
# GPU suite + Lagrangian-path timing + one bench line.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_r02p.sh r02p'
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r02p}
mkdir -p "$O"
( time timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --durations=8 ) > "$O/pytest_gpu.log" 2>&1; tail -14 "$O/pytest_gpu.log"
python tools/lagrangian_time.py 256 exp > "$O/lagrangian.txt" 2>&1; python tools/lagrangian_time.py 256 tophat >> "$O/lagrangian.txt" 2>&1; cat "$O/lagrangian.txt" | tail -6
python bench.py --steps 10 --warmup 3 --no-yardstick --no-cpu-baseline > "$O/bench.json" 2> "$O/bench.err"; cut -c1-400 "$O/bench.json"
