#!/bin/bash
# ncu --set full on kernels matching a regex inside the default bench; usage: gpu_ncu_k.sh <tag> <regex> <skip> <count>
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/$1
mkdir -p "$O"
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"$2" -s ${3:-45} -c ${4:-4} -o "$O/prof" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$O/ncu.log" 2>&1
tail -3 "$O/ncu.log" | cut -c1-300; ls -la "$O"
