#!/bin/bash
# Round 2, call d: single-sweep ladder with per-CTA queue segments -- bit-identity tests, A/B bench, e2e split,
# ncu of the sweep kernels.   usage: gpurun --timeout 1200 -- 'bash tools/gpu_r02d.sh r02d'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r02d}
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi -L > "$O/gpu.txt"
( time timeout 600 python -m pytest tests/test_single_sweep.py -m gpu -q -x --durations=5 ) > "$O/pytest_sweep.log" 2>&1; tail -8 "$O/pytest_sweep.log"
B200_TIMING=1 python bench.py --steps 10 --warmup 3 --no-yardstick --no-cpu-baseline > "$O/bench.json" 2> "$O/bench.err"; grep -m1 speculation "$O/bench.err" | cut -c1-600
B200_SPEC=0 python bench.py --steps 10 --warmup 3 --no-yardstick --no-cpu-baseline --no-e2e > "$O/bench_nospec.json" 2> "$O/bench_nospec.err"
python - "$O" <<'PY'
import json, sys
o = sys.argv[1]
for f in ("bench.json", "bench_nospec.json"):
    try:
        d = json.loads(open(f"{o}/{f}").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["config"]["ms_perturb"], d["config"]["ms_ionize"], d["config"]["global_xH"], {k: round(v, 3) for k, v in d["kernel_profile_ms_per_step"].items()})
        print(d.get("e2e"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'fcoll_sum_kernel|spec_resolve_kernel' -s 20 -c 4 -o "$O/prof_spec" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-yardstick > "$O/ncu.log" 2>&1
tail -2 "$O/ncu.log" | cut -c1-300; ls -la "$O"
