cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --hii-dim 256 > gpurun_out/bench_256g.json 2> gpurun_out/bench_256g.err; tail -2 gpurun_out/bench_256g.err
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --hii-dim 512 --dim 1024 --box-len 768 --r-bubble-max 40 > gpurun_out/bench_512d.json 2> gpurun_out/bench_512d.err; tail -2 gpurun_out/bench_512d.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_256g.json','gpurun_out/bench_512d.json'):
    d=json.load(open(f))
    print(f, d['value'], d['ms_per_step'], d['config']['ms_perturb'], d['config']['ms_ionize'])
    print(' ', d.get('kernel_profile_ms_per_step'))
    print(' ', d.get('roofline'))
PY
nvidia-smi --query-gpu=name,memory.total --format=csv; free -g | head -2; nproc
