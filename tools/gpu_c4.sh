#!/bin/bash
# usage: gpurun --gpus 8 --timeout 150 -- 'bash tools/gpu_c4.sh r02s'
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/${1:-r02s}
mkdir -p "$O"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
    tools/c4_slab_pipeline.py --hii-dim 1024 --dim 2048 --box-len 1000 --reps 2 > "$O/c4.json" 2> "$O/c4.err"
echo "rc=$?"; tail -3 "$O/c4.err" | cut -c1-400; cat "$O/c4.json" | cut -c1-1500
