#!/bin/bash
# Round 2, call b (2 GPUs): slab-decomposed box tests, the rest of the GPU suite (512 parity test), bench with
# both strong partitions.   usage: gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r02b.sh r02b'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r02b}
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi -L > "$O/gpu.txt"
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --durations=5 ) > "$O/pytest_multi.log" 2>&1; tail -30 "$O/pytest_multi.log"
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=8 -k "c2 or 512 or device_entry" -s ) > "$O/pytest_parity.log" 2>&1; tail -25 "$O/pytest_parity.log"
python bench.py --steps 5 --warmup 3 --no-yardstick > "$O/bench1.json" 2> "$O/bench1.err"; tail -2 "$O/bench1.err"
python - "$O/bench1.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=1", d["ms_per_step"], d["config"]["ms_perturb"], d["config"]["ms_ionize"], {k: round(v, 3) for k, v in d["kernel_profile_ms_per_step"].items()})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 \
    bench.py --gpus 2 --steps 5 --warmup 3 > "$O/bench2.json" 2> "$O/bench2.err"; tail -5 "$O/bench2.err"; cat "$O/bench2.json"
B200_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 \
    bench.py --gpus 2 --steps 5 --warmup 3 --partition slab > "$O/bench2_slab.json" 2> "$O/bench2_slab.err"; tail -5 "$O/bench2_slab.err"; cat "$O/bench2_slab.json"
