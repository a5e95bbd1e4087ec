#!/bin/bash
# profiles for the default workload: launch list of one timed step + ncu --set full of the hot kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/$1
mkdir -p "$O"
# IC generation launches ~170 kernels, the warm-up step ~340: skip them, list the timed step only
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$O/launches.csv" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$O/ncu_launch.log" 2>&1
tail -2 "$O/ncu_launch.log" | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'fft_strided_pow2_kernel<.int.512|fft_c2r_z_pow2_kernel<.int.256|fft_r2c_z_pow2_kernel<.int.256|fcoll_sum_kernel|ionise_delta_kernel|move_cic_grouped|window_expand' -s ${2:-262} -c ${3:-14} \
    -o "$O/prof" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$O/ncu_full.log" 2>&1
tail -2 "$O/ncu_full.log" | cut -c1-200; ls -la "$O"
