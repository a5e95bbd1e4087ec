cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_256b.json 2> gpurun_out/bench_256b.err; tail -3 gpurun_out/bench_256b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_256b.json'))
for k in ('value','ms_per_step','e2e','cpu_baseline','roofline','step_roofline','kernel_profile_ms_per_step','gpu_launches'):
    print(k, d.get(k))
print(d['config'])
PY
