#!/bin/bash
# Round 2, call c: the single-sweep (speculative) ladder -- parity at C2 / 512 / device entry, A/B bench against
# B200_SPEC=0, launch list, and a full ncu capture of the new kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_r02c.sh r02c'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r02c}
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi -L > "$O/gpu.txt"
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=8 -s ) > "$O/pytest_parity.log" 2>&1; tail -12 "$O/pytest_parity.log"
B200_TIMING=1 python bench.py --steps 10 --warmup 3 --no-yardstick --no-cpu-baseline > "$O/bench.json" 2> "$O/bench.err"; grep -m2 speculation "$O/bench.err" | cut -c1-1500
B200_SPEC=0 python bench.py --steps 10 --warmup 3 --no-yardstick --no-cpu-baseline --no-e2e > "$O/bench_nospec.json" 2> "$O/bench_nospec.err"
python - "$O" <<'PY'
import json, sys
o = sys.argv[1]
for f in ("bench.json", "bench_nospec.json"):
    try:
        d = json.loads(open(f"{o}/{f}").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["config"]["ms_perturb"], d["config"]["ms_ionize"], d["config"]["global_xH"], {k: round(v, 3) for k, v in d["kernel_profile_ms_per_step"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'fcoll_sum_kernel|spec_resolve_kernel|move_cic_grouped' -s 20 -c 5 -o "$O/prof_spec" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-yardstick > "$O/ncu.log" 2>&1
tail -2 "$O/ncu.log" | cut -c1-300; ls -la "$O"
