import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import common
pkg = common.pkg
be = pkg.get_backend(); be.set_table_path(common.table_dir())
os.environ["B200_SKIP_SCRATCH_OUTPUTS"] = "1"
for hii, dim in ((256, 768), (512, 1024), (512, 1536)):
    inputs = common.make_inputs(hii=hii, dim=dim, box_len=1.5 * hii)
    for mode in ("device", "host"):
        if mode == "device": os.environ["B200_IC_RNG"] = "device"
        else: os.environ.pop("B200_IC_RNG", None)
        t0 = time.perf_counter()
        ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
        t1 = time.perf_counter()
        import ctypes as C
        ms = C.c_double(); be.lib.b200_last_call_stats(None, None, None, C.byref(ms))
        print(f"ICs HII={hii} DIM={dim} rng={mode}: wall {t1-t0:.2f} s, device timer {ms.value/1e3:.2f} s", flush=True)
        del ics
