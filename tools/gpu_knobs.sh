#!/bin/bash
# A/B of run-time knobs of the ladder in one call (device leg only).  usage: gpurun --timeout 400 -- 'bash tools/gpu_knobs.sh r02u'
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r02u}
mkdir -p "$O"
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-yardstick > "$O/$name.json" 2> "$O/$name.err"
  python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["kernel_profile_ms_per_step"]
    print(sys.argv[2], round(d["ms_per_step"], 2), round(d["config"]["ms_perturb"], 2), round(d["config"]["ms_ionize"], 2), d["config"]["global_xH"],
          {n: round(k.get(n, 0), 2) for n in ("fcoll_sum_classify_kernel", "spec_resolve_kernel", "fft_strided_pow2_kernel", "fft_c2r_z_pow2_kernel")})
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
}
run base B200_NOP=1
run ahead2 B200_IONIZE_AHEAD=2
run eps01 B200_SPEC_EPS=0.01
run ctas4 B200_SWEEP_CTAS=4
run ctas6 B200_SWEEP_CTAS=6
run notabovl B200_TABLE_OVERLAP=0
