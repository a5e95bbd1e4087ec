#!/bin/bash
# 2 GPUs: multi-GPU tests (slab ICs, slab box), CLASS-table tests, scaling bench at N=2.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r02i}
mkdir -p "$O"
nvidia-smi -L > "$O/gpu.txt"; nproc >> "$O/gpu.txt"
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_class_tables.py -m gpu -q -x --durations=6 ) > "$O/pytest.log" 2>&1; tail -14 "$O/pytest.log"
bash tools/gpu_scale.sh ${1:-r02i} 2
