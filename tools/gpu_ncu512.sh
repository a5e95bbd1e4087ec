#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-n}
O=gpurun_out/$TAG
mkdir -p "$O"
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'fft_strided_pow2_kernel<.int.512|fft_c2r_z_pow2_kernel<.int.256|fcoll_sum_kernel|ionise_delta_kernel|move_cic_grouped' -s ${2:-215} -c ${3:-12} \
    -o "$O/prof_512" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --hii-dim 512 --dim 1024 --box-len 768 --r-bubble-max 40 > "$O/ncu_full.log" 2>&1
tail -5 "$O/ncu_full.log"; ls -la "$O"
