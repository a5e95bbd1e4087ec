#!/usr/bin/env python
"""Wall / device time of the Lagrangian-source path on the GPU (ComputeHaloBox + ComputeIonizedBox with a HaloBox),
host-pointer entry points.   python tools/lagrangian_time.py [HII_DIM] [exp|tophat]"""
import ctypes as C
import dataclasses
import importlib
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import common  # noqa: E402

pkg = importlib.import_module("21cmfast_b200")
hii = int(sys.argv[1]) if len(sys.argv) > 1 else 256
exp_filter = (sys.argv[2] if len(sys.argv) > 2 else "exp") == "exp"
be = common.gpu_backend()
os.environ["B200_IC_RNG"] = "device"
inp = common.make_inputs(hii=hii, dim=2 * hii, box_len=1.5 * hii, source="L-INTEGRAL", R_BUBBLE_MAX=40.0)
ao = dataclasses.replace(inp.astro_options, USE_EXP_FILTER=exp_filter, CELL_RECOMB=exp_filter)
inputs = dataclasses.replace(inp, astro_options=ao)
ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
pf = pkg.perturb_field(redshift=7.0, initial_conditions=ics, backend=be)


def ms():
    a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_double()
    be.lib.b200_last_call_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    return d.value


be.lib.b200_profile_report.argtypes = [C.c_char_p, C.c_int]
for rep in range(3):
    if rep == 2:
        be.lib.b200_profile_enable(1)
    t0 = time.perf_counter()
    hb = pkg.compute_halobox(redshift=7.0, initial_conditions=ics, backend=be)
    t1 = time.perf_counter()
    ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, halobox=hb, backend=be)
    t2 = time.perf_counter()
    print(f"HII_DIM={hii} {'exp-mfp' if exp_filter else 'top-hat'} filter: halobox {1e3 * (t1 - t0):.1f} ms wall, "
          f"ionize {1e3 * (t2 - t1):.1f} ms wall ({ms():.1f} ms device incl. copies), xH={ib.global_xH:.4f}")
buf = C.create_string_buffer(1 << 16)
be.lib.b200_profile_report(buf, len(buf))
print(buf.value.decode())
