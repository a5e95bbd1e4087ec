#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench at the C2/C3 sizes, ncu launch list + full capture.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_check.sh <tag>'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > "$O/box.txt"; free -g | head -2 >> "$O/box.txt"; nproc >> "$O/box.txt"
( time python -m pytest tests -m gpu -q -x ) > "$O/pytest_gpu.log" 2>&1; tail -4 "$O/pytest_gpu.log"
python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -2 "$O/smoke.log"
python bench.py --steps 5 --warmup 3 > "$O/bench_512.json" 2> "$O/bench_512.err"; tail -2 "$O/bench_512.err"
python bench.py --steps 5 --warmup 3 --hii-dim 256 --box-len 300 --r-bubble-max 15 --ref-hii-dim 128 --no-cpu-baseline > "$O/bench_256.json" 2> "$O/bench_256.err"; tail -2 "$O/bench_256.err"
python bench.py --impl reference --steps 1 --warmup 1 > "$O/bench_ref.json" 2> "$O/bench_ref.err"; tail -2 "$O/bench_ref.err"
# launch list (cold-cache, serialised): one warm-up + one timed step at the default workload
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$O/launches_256.csv" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$O/ncu_launch.log" 2>&1
# full capture of the FFT passes + sweeps (3 launches each, after the warm-up step)
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'fft_strided_pow2|fft_c2r_z_pow2|fft_r2c_z_pow2|fcoll_sum|ionise_kernel|move_cic' -s 40 -c 24 \
    -o "$O/prof_256" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$O/ncu_full.log" 2>&1
python - <<PY
import json
for f in ('bench_256','bench_512','bench_ref'):
    try:
        d=json.load(open('$O/'+f+'.json'))
        print(f, d.get('value'), d.get('ms_per_step'), d.get('e2e'), d.get('roofline'))
        print(' ', d.get('kernel_profile_ms_per_step'))
        print(' ', d.get('cpu_baseline'), d.get('step_roofline'))
    except Exception as e:
        print(f, 'ERR', e)
PY
ls -la "$O"
