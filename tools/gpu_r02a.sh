#!/bin/bash
# Round 2, first GPU call: the whole GPU suite (incl. the new C2 / 512 / device-entry / cffi tests),
# smoke, both bench arms at the default workload, launch list.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_r02a.sh r02a'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r02a}
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi --query-gpu=name,memory.total --format=csv > "$O/gpu.txt"; free -g | head -2 >> "$O/gpu.txt"; nproc >> "$O/gpu.txt"
( time python -m pytest tests -m gpu -q -x --durations=15 ) > "$O/pytest_gpu.log" 2>&1; tail -25 "$O/pytest_gpu.log"
python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
python bench.py --steps 10 --warmup 3 > "$O/bench.json" 2> "$O/bench.err"; tail -2 "$O/bench.err"; cat "$O/bench.json"
( time python bench.py --impl reference --steps 20 --warmup 5 ) > "$O/bench_ref.json" 2> "$O/bench_ref.err"; tail -4 "$O/bench_ref.err"; cat "$O/bench_ref.json"
B200_CIC_DOUBLE=1 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-yardstick > "$O/bench_cic_double.json" 2>> "$O/bench.err"
python - <<'PY' "$O"
import json, sys
o = sys.argv[1]
for f in ("bench.json", "bench_cic_double.json"):
    try:
        d = json.loads(open(f"{o}/{f}").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["config"]["ms_perturb"], d["config"]["ms_ionize"], d.get("kernel_profile_ms_per_step", {}).get("move_cic_grouped_kernel"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 600 --csv --log-file "$O/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-yardstick > "$O/ncu_bench.log" 2>&1
python tools/launch_summary.py "$O/launches.csv" > "$O/launches.md" 2>/dev/null; head -30 "$O/launches.md"
