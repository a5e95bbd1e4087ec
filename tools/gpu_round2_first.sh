#!/bin/bash
# First GPU call of the next round: full parity suite, default bench, and timings of the paths that were
# built after the round-1 GPU budget ran out of profiling time (recombination / x_e route, sphere painting,
# N_THREADS > 1 ICs), so that their kernels get a launch list before anyone tunes them.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r02a'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r02a}
O=gpurun_out/$TAG
mkdir -p "$O"
( time python -m pytest tests -m gpu -q -x ) > "$O/pytest_gpu.log" 2>&1; tail -4 "$O/pytest_gpu.log"
python __graft_entry__.py smoke > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
python - > "$O/gamma_approx_gpu.txt" 2>&1 <<'PY'
# option-matrix cases that so far only ran on the CPU tier: run them on the GPU, then move them into GPU_CASES
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import common, test_option_matrix as m
for name in m.CASES:
    if name not in m.GPU_CASES:
        m._run_case(common.gpu_backend(), common.ref_backend(), name)
        print("ok", name)
PY
cat "$O/gamma_approx_gpu.txt"
python bench.py --steps 5 --warmup 3 > "$O/bench.json" 2> "$O/bench.err"; tail -2 "$O/bench.err"
python - > "$O/evolution_timing.txt" 2>&1 <<'PY'
import sys, time, dataclasses
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, common
import test_recombinations as tr
pkg = common.pkg
be = common.gpu_backend()
for hii in (128, 256):
    for name in ("inhomogeneous_filtered", "ts_fluct_inhomogeneous_filtered"):
        inputs = tr._inputs(hii=hii, **tr.CASES[name])
        ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
        pfs = [pkg.perturb_field(redshift=z, initial_conditions=ics, backend=be) for z in tr.REDSHIFTS]
        t = time.time(); out = tr._chain(be, inputs, ics, pfs); dt = time.time() - t
        print(f"{name} HII_DIM={hii}: {dt / len(pfs) * 1e3:.1f} ms per ionize call (host buffers), xH={out[-1].global_xH:.4f}")
    import test_sphere_painting as sp
    inputs = sp._inputs("E-INTEGRAL", hii, 1.0)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=be)
    t = time.time(); ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be); dt = time.time() - t
    print(f"sphere painting HII_DIM={hii}: {dt * 1e3:.1f} ms, xH={ib.global_xH:.4f}")
for nt in (1, 16):
    inputs = common.make_inputs(hii=256, dim=768, seed=1, n_threads=nt)
    t = time.time(); pkg.compute_initial_conditions(inputs=inputs, backend=be); print(f"ICs DIM=768 N_THREADS={nt}: {time.time() - t:.2f} s")
PY
cat "$O/evolution_timing.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$O/launches_evolution.csv" \
    python -m pytest tests/test_recombinations.py tests/test_sphere_painting.py -m gpu -q -x -k "inhomogeneous_filtered or sphere" > "$O/ncu_evolution.log" 2>&1
python tools/launch_summary.py "$O/launches_evolution.csv" > "$O/launches_evolution.md" 2>/dev/null; head -20 "$O/launches_evolution.md"
