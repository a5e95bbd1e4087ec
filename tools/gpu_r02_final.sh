#!/bin/bash
# Round-2 evidence run on ONE B200: whole GPU suite, smoke, both bench arms at the default workload, other sizes,
# ncu launch list and a full capture of the hot kernels from the final code.
# usage: gpurun --timeout 2700 -- 'bash tools/gpu_r02_final.sh r02_final'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r02_final}
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi --query-gpu=name,memory.total --format=csv > "$O/gpu.txt"; nproc >> "$O/gpu.txt"
python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
python bench.py --steps 10 --warmup 3 > "$O/bench.json" 2> "$O/bench.err"; tail -2 "$O/bench.err" | cut -c1-300; cut -c1-600 "$O/bench.json"
B200_SPEC=0 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-yardstick > "$O/bench_two_sweeps.json" 2>> "$O/bench.err"
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-yardstick --hii-dim 256 --dim 768 --box-len 300 --r-bubble-max 15 > "$O/bench_256.json" 2>> "$O/bench.err"
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-yardstick --hii-dim 1024 --dim 1024 --box-len 1000 --r-bubble-max 15 > "$O/bench_1024.json" 2>> "$O/bench.err"
for f in bench_two_sweeps bench_256 bench_1024; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["ms_per_step"], 2), "%.3g cells/s" % d["value"], round(d["config"]["ms_perturb"], 2), round(d["config"]["ms_ionize"], 2), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file "$O/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-yardstick > "$O/ncu_launches.log" 2>&1
python tools/launch_summary.py "$O/launches.csv" > "$O/launches.md" 2>/dev/null; head -24 "$O/launches.md"
# two small full captures (the reports are too large to travel: keep the raw page and a summary, drop the .ncu-rep)
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'move_cic_grouped|ionise_last_fused' -s 0 -c 2 \
    -o "$O/prof_a" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-yardstick > "$O/ncu_a.log" 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled \
    -k regex:'fft_strided_pow2_kernel|fft_c2r_z_pow2_kernel|fcoll_sum_kernel|spec_resolve_kernel' -s 22 -c 14 \
    -o "$O/prof_b" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-yardstick > "$O/ncu_b.log" 2>&1
tail -1 "$O/ncu_a.log" | cut -c1-200; tail -1 "$O/ncu_b.log" | cut -c1-200
for r in a b; do
  ncu -i "$O/prof_$r.ncu-rep" --page raw --csv > "$O/ncu_full_raw_$r.csv" 2>/dev/null
  python tools/ncu_summary.py "$O/prof_$r.ncu-rep" "$O/ncu_full_$r" --hii 512 > "$O/ncu_full_$r.md" 2>&1; cat "$O/ncu_full_$r.md"
  rm -f "$O/prof_$r.ncu-rep"
done
( time python bench.py --impl reference --steps 20 --warmup 5 ) > "$O/bench_ref.json" 2> "$O/bench_ref.err"; tail -3 "$O/bench_ref.err"; cut -c1-500 "$O/bench_ref.json"
ls -la "$O"
