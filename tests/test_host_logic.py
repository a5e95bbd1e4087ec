"""Host-side logic on the CPU tier: the Python mirror of the reference's input structs, the host
scalar chain of the *shipped* library (g++-compiled units of lib21cmfast_b200.so, no device
work) against golden values produced by the compiled reference, and the full pipeline of the
host-emulation build (tests/_emu, same kernel sources, one logical thread per block) against the
golden fixtures."""
import ctypes as C
import math
import os
from pathlib import Path

import numpy as np
import pytest

import common

pkg = common.pkg
ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "21cmfast_b200" / "csrc" / "lib21cmfast_b200.so"


def test_input_defaults_and_transformers():
    ap = pkg.AstroParams()
    c = ap.cdict
    assert math.isclose(c["F_STAR10"], 10**-1.3) and math.isclose(c["M_TURN"], 10**8.7)
    assert math.isclose(c["F_STAR7_MINI"], 10 ** (-1.3 - 1.5)) and math.isclose(c["L_X_MINI"], 10**40.5)
    assert math.isclose(c["SIGMA_STAR"], 0.25 * math.log(10))
    so = pkg.SimulationOptions(HII_DIM=50)
    assert so.dim == 150 and math.isclose(so.box_len, 75.0)
    with pytest.raises(ValueError):
        pkg.SimulationOptions(HII_DIM=50, DIM=100, HIRES_TO_LOWRES_FACTOR=2)
    with pytest.raises(ValueError):
        pkg.MatterOptions(FILTER="sharp-k")
    mo = pkg.MatterOptions(SOURCE_MODEL="E-INTEGRAL")
    assert mo.cdict["SOURCE_MODEL"] == 1 and mo.cdict["POWER_SPECTRUM"] == 0 and mo.cdict["HMF"] == 1
    cp = pkg.CosmoParams()
    assert math.isclose(cp.OMl, 1 - cp.OMm)
    inp = common.make_inputs()
    assert not inp.evolution_required
    assert inp.evolve_input_structs(HII_DIM=16).simulation_options.HII_DIM == 16


def test_output_struct_allocation_rules():
    inp = common.make_inputs(hii=8, dim=16)
    ics = pkg.InitialConditions.new(inp)
    assert ics.hires_vx is None and ics.lowres_vx.shape == (8, 8, 8) and ics.hires_vx_2LPT.shape == (16, 16, 16)
    ib = pkg.IonizedBox.new(inp, 8.0)
    assert (ib.neutral_fraction == 1).all() and ib.unnormalised_nion.shape == (1, 8, 8, 8)
    assert ib.cumulative_recombinations is None
    s = ib.cstruct
    assert not s.cumulative_recombinations and bool(s.neutral_fraction)


def _host_backend():
    if not LIB.exists():
        pytest.skip("CUDA library not built")
    be = pkg.Backend()
    be.set_table_path(common.table_dir())
    return be


def test_shipped_library_host_scalars_match_reference_golden():
    be = _host_backend()
    g = np.load(common.GOLDEN / "host_scalars.npz")
    be.state.init(common.make_inputs(), broadcast_inputs=True, ps=True, sigma=True, heat=True)
    for z, d in zip(g["z"], g["dicke"]):
        assert abs(be.lib.dicke(float(z)) - d) <= 1e-14 * abs(d)
    for M, s, ds in zip(g["M"], g["sigma"], g["dsigmasqdm"]):
        assert abs(be.lib.sigma_z0(float(M)) - s) <= 1e-10 * abs(s)
        assert abs(be.lib.dsigmasqdm_z0(float(M)) - ds) <= 1e-8 * abs(ds)
    for k, p in zip(g["k"], g["power"]):
        assert abs(be.lib.power_in_k(float(k)) - p) <= 1e-12 * abs(p)


@pytest.mark.parametrize("case", list(common.GOLDEN_CASES))
def test_shipped_library_ionize_host_chain_matches_golden(case):
    """mean_f_coll (QAG-61 over the mass function) from the shipped library's host code equals the
    value the compiled reference produced, to 1e-12 -- the check that caught nvcc's float-overload
    host pass."""
    be = _host_backend()
    inputs, _, _, g_ib = common.load_golden(case)
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True)
    out = (C.c_double * 80)()
    assert be.lib.b200_ionize_host_scalars(C.c_float(8.0), out, 80) == 0
    assert abs(out[0] - g_ib.mean_f_coll) <= 1e-12 * g_ib.mean_f_coll
    n_radii = int(out[7])
    radii = np.array([out[8 + 2 * i] for i in range(n_radii)])
    assert n_radii == int(math.log(min(15.0, 0.620350491 * 48) / max(0.620350491, 0.620350491 * 1.5)) / math.log(1.1) + 1)
    np.testing.assert_allclose(radii[1:] / radii[:-1], 1.1, rtol=1e-12)


@pytest.mark.parametrize("case", list(common.GOLDEN_CASES))
def test_emulated_kernels_reproduce_golden(case):
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built (run __graft_entry__.build())")
    inputs, ics, g_pf, g_ib = common.load_golden(case)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
    common.compare_struct(pf, g_pf, tols={"velocity_z": common.TOL_VELOCITY})
    ib = pkg.compute_ionization_field(perturbed_field=g_pf, initial_conditions=ics, backend=emu)
    assert common.compare_ionized(ib, g_ib)["mask_mismatch"] == 0


def test_emulated_initial_conditions_reproduce_golden_seed_stream():
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    inputs, g_ics, _, _ = common.load_golden(common.GOLDEN_BASE)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
    common.compare_struct(ics, g_ics)
    # feeding hires_density back in reproduces every other field (reference
    # tests/test_initial_conditions.py:153-178)
    again = pkg.compute_initial_conditions(inputs=inputs, backend=emu, initial_density=ics.hires_density)
    common.compare_struct(again, ics, tol=1e-5, skip=("hires_density", "hires_vx_2LPT", "hires_vy_2LPT", "hires_vz_2LPT"))


def test_emulated_generic_cic_equals_grouped():
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    inputs, ics, _, _ = common.load_golden(common.GOLDEN_BASE)
    a = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
    os.environ["B200_CIC_GENERIC"] = "1"
    try:
        b = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
    finally:
        os.environ.pop("B200_CIC_GENERIC")
    assert np.abs(a.density - b.density).max() <= 2e-7 * np.abs(a.density).max() + 1e-7


def test_error_codes_follow_reference_convention():
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    inputs = common.make_inputs(hii=8, dim=16)
    bad = inputs.clone(astro_options=pkg.AstroOptions(USE_EXP_FILTER=False, CELL_RECOMB=False, USE_TS_FLUCT=True))
    ics = pkg.InitialConditions.new(inputs)
    pf = pkg.PerturbedField.new(bad, 8.0)
    emu.state.init(bad, broadcast_inputs=True, ps=True, sigma=True, heat=True)
    ib = pkg.IonizedBox.new(bad, 8.0)
    prev = pkg.IonizedBox.initial(bad)
    ts, hb = pkg.outputs.TsBox.dummy(bad), pkg.outputs.HaloBox.dummy(bad)
    st = emu.lib.ComputeIonizedBox(C.c_float(8.0), C.c_float(-1.0), C.byref(pf.cstruct), C.byref(pf.cstruct),
                                   C.byref(prev.cstruct), C.byref(ts.cstruct), C.byref(hb.cstruct),
                                   C.byref(ics.cstruct), C.byref(ib.cstruct))
    assert st == 3  # ValueError: outside the scoped path, reported through the status code
    assert (prev.z_reion == -1).all()  # first-snapshot side effect on the previous box is preserved


def test_emulated_smoothed_perturb_matches_reference():
    """SMOOTH_EVOLVED_DENSITY_FIELD through the host-emulated kernels against the compiled reference
    (CPU tier twin of tests/test_gpu_parity.py::test_smoothed_evolved_density_vs_reference)."""
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    inputs = common.make_inputs(hii=16, dim=32, smooth_evolved=True)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    r_pf = pkg.perturb_field(redshift=7.0, initial_conditions=ics, backend=ref)
    pf = pkg.perturb_field(redshift=7.0, initial_conditions=ics, backend=emu)
    common.compare_struct(pf, r_pf, tols={"velocity_z": common.TOL_VELOCITY})


def _np_gaussians_from_raw(raw, want):
    """GSL rules in numpy/python: uniform_pos redraws zero words, polar Box-Muller keeps y."""
    out, p = [], 0
    def upos():
        nonlocal p
        while True:
            w = int(raw[p]); p += 1
            if w != 0:
                return w / 4294967296.0
    while len(out) < want:
        x = -1 + 2 * upos(); y = -1 + 2 * upos()
        r2 = x * x + y * y
        if r2 > 1.0 or r2 == 0:
            continue
        out.append(y * math.sqrt(-2.0 * math.log(r2) / r2))
    return np.array(out), p


def test_device_gaussian_stream_equals_sequential_generator():
    """csrc/gslrng.cu (three-step parallel MT19937 refresh, data-parallel polar attempts, prefix-sum
    placement) through the host-emulated kernels == the sequential generator, bit for bit, across
    refresh boundaries, chunk boundaries and carried-over words; and the MT19937 words themselves
    against numpy's legacy seeding."""
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    lib = emu.lib
    lib.b200_gsl_gaussian_stream.argtypes = [C.c_ulong, C.c_longlong, C.c_longlong, C.POINTER(C.c_double)]
    lib.b200_host_gaussian_stream.argtypes = [C.c_ulong, C.c_longlong, C.POINTER(C.c_double)]
    for seed, n1, n2 in ((12345, 1000, 3), (1, 5, 70001), (4357, 20000, 20000), (987654321, 1, 1)):
        a = np.zeros(n1 + n2); b = np.zeros(n1 + n2)
        assert lib.b200_gsl_gaussian_stream(seed, n1, n2, a.ctypes.data_as(C.POINTER(C.c_double))) == 0
        assert lib.b200_host_gaussian_stream(seed, n1 + n2, b.ctypes.data_as(C.POINTER(C.c_double))) == 0
        assert np.array_equal(a, b), (seed, n1, n2, np.flatnonzero(a != b)[:5])
    # the sequential generator itself against numpy's MT19937 (legacy seeding = init_genrand)
    rs = np.random.RandomState(12345)
    raw = rs.randint(0, 2**32, size=4000, dtype=np.uint64).astype(np.uint32)
    exp, _ = _np_gaussians_from_raw(raw, 1000)
    got = np.zeros(1000)
    lib.b200_host_gaussian_stream(12345, 1000, got.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.allclose(got, exp, rtol=0, atol=1e-15)


def test_gaussians_from_raw_redraws_zero_words():
    """The sequential fallback used for chunks that contain a zero word (GSL's uniform_pos redraws
    it, which shifts the pairing of all later words)."""
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    lib = emu.lib
    lib.b200_gaussians_from_raw_host.argtypes = [C.POINTER(C.c_uint), C.c_longlong, C.c_longlong, C.POINTER(C.c_double)]
    lib.b200_gaussians_from_raw_host.restype = C.c_longlong
    rng = np.random.default_rng(3)
    raw = rng.integers(1, 2**32, size=5000, dtype=np.uint64).astype(np.uint32)
    raw[[0, 7, 8, 1001, 2500]] = 0
    exp, used = _np_gaussians_from_raw(raw, 1500)
    got = np.zeros(1500)
    n = lib.b200_gaussians_from_raw_host(raw.ctypes.data_as(C.POINTER(C.c_uint)), len(raw), 1500,
                                         got.ctypes.data_as(C.POINTER(C.c_double)))
    assert n == used
    assert np.allclose(got, exp, rtol=0, atol=1e-15)
    assert lib.b200_gaussians_from_raw_host(raw.ctypes.data_as(C.POINTER(C.c_uint)), 100, 1500,
                                            got.ctypes.data_as(C.POINTER(C.c_double))) == -1


def test_bench_reference_arm_line_contract():
    """bench.py --impl reference prints ONE JSON line with the keys the driver reads (metric, unit,
    config identical to the b200 arm's, impl, cpu_baseline, e2e with zero copy bytes)."""
    import json
    import subprocess
    import sys
    if common.ref_backend() is None:
        pytest.skip("needs oracle/_ref")
    root = Path(__file__).resolve().parent.parent
    # a small grid here (the CPU tier has minutes, the default workload takes 44 s per step on 16 cores); the arm
    # runs the workload it prints, so the label must carry these sizes -- and the defaults the benchmark's
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--hii-dim", "32", "--dim", "96", "--box-len", "48", "--r-bubble-max", "10"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("coeval cells/sec")
    assert "HII_DIM=32 DIM=96 BOX_LEN=48" in d["config"]["workload"]
    assert "HII_DIM=32 DIM=96" in d["cpu_baseline"]["sample"]          # what ran is what the line says
    assert abs(d["value"] - 32**3 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod_ref", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    old_argv, sys.argv = sys.argv, ["bench.py"]
    try:
        dflt = bench.parse()
    finally:
        sys.argv = old_argv
    assert bench.workload(dflt) == (512, 1536, 768.0) and bench.n_radii(512, 768.0, dflt.r_bubble_max) == 40
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]


def test_bench_radius_count_matches_library():
    """bench.py's n_radii() (used for the algorithmic-byte count) equals the ladder the library builds."""
    import importlib.util
    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("bench_mod", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    for hii, box, rmax in ((512, 768.0, 40.0), (256, 300.0, 15.0), (1024, 1000.0, 15.0), (64, 96.0, 15.0)):
        inputs = common.make_inputs(hii=hii, dim=hii, box_len=box, R_BUBBLE_MAX=rmax)
        emu.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True)
        out = (C.c_double * 8)()
        emu.lib.b200_ionize_host_scalars.argtypes = [C.c_float, C.POINTER(C.c_double), C.c_int]
        assert emu.lib.b200_ionize_host_scalars(C.c_float(8.0), out, 8) == 0
        assert int(out[7]) == bench.n_radii(hii, box, rmax), (hii, box, rmax, out[7])
