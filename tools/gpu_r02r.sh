#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r02r}
mkdir -p "$O"
( time timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --durations=5 ) > "$O/pytest_gpu.log" 2>&1; tail -10 "$O/pytest_gpu.log"
python bench.py --steps 10 --warmup 3 --no-yardstick --no-cpu-baseline > "$O/bench.json" 2> "$O/bench.err"
python - "$O/bench.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 2), round(d["config"]["ms_perturb"], 2), round(d["config"]["ms_ionize"], 2), d["config"]["global_xH"], {k: round(v, 2) for k, v in d["kernel_profile_ms_per_step"].items()}, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["step_roofline"]["frac"])
PY
