"""GPU parity tests: the CUDA library (through the C-ABI, host buffers) against the compiled
reference in oracle/_ref when it travelled with the repo, and against the committed golden
fixtures always.  Tolerances are the float32 bounds of tests/common.py."""
import ctypes as C

import numpy as np
import pytest

import common

pkg = common.pkg
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    return common.gpu_backend()


@pytest.mark.parametrize("case", list(common.GOLDEN_CASES))
def test_golden_perturb_ionize(gpu, case):
    """perturb + ionize on the golden ICs reproduce the reference outputs (golden fixtures)."""
    inputs, ics, g_pf, g_ib = common.load_golden(case)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=gpu)
    common.compare_struct(pf, g_pf, tols={"velocity_z": common.TOL_VELOCITY})
    ib = pkg.compute_ionization_field(perturbed_field=g_pf, initial_conditions=ics, backend=gpu)
    stats = common.compare_ionized(ib, g_ib)
    assert stats["mask_mismatch"] == 0


def test_golden_initial_conditions(gpu):
    """ICs from the seed reproduce the reference's field (GSL MT19937 stream parity, N_THREADS=1)."""
    inputs, g_ics, _, _ = common.load_golden(common.GOLDEN_BASE)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=gpu)
    common.compare_struct(ics, g_ics)


@pytest.mark.parametrize("hii,dim,source,filt,perturb", [
    (64, 128, "E-INTEGRAL", "spherical-tophat", "2LPT"),
    (128, 256, "E-INTEGRAL", "spherical-tophat", "2LPT"),   # N = 128 / 256 register FFTs, window-table x pass
    (64, 128, "CONST-ION-EFF", "sharp-k", "2LPT"),
    (35, 70, "CONST-ION-EFF", "spherical-tophat", "ZELDOVICH"),
    (50, 150, "E-INTEGRAL", "gaussian", "2LPT"),
    (48, 96, "E-INTEGRAL", "spherical-tophat", "LINEAR"),
])
def test_pipeline_vs_reference(gpu, hii, dim, source, filt, perturb):
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box (golden tests cover parity)")
    inputs = common.make_inputs(hii=hii, dim=dim, source=source, hii_filter=filt, perturb=perturb)
    r_ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=gpu)
    common.compare_struct(ics, r_ics)
    r_pf = pkg.perturb_field(redshift=7.5, initial_conditions=r_ics, backend=ref)
    pf = pkg.perturb_field(redshift=7.5, initial_conditions=r_ics, backend=gpu)
    common.compare_struct(pf, r_pf, tols={"velocity_z": common.TOL_VELOCITY})
    r_ib = pkg.compute_ionization_field(perturbed_field=r_pf, initial_conditions=r_ics, backend=ref)
    ib = pkg.compute_ionization_field(perturbed_field=r_pf, initial_conditions=r_ics, backend=gpu)
    common.compare_ionized(ib, r_ib)


@pytest.mark.parametrize("filter_flag", [0, 1])
def test_filter_512_vs_reference(gpu, filter_flag):
    """The N = 512 kernels of the metric's grid size (three radix-8 stages, line-per-warp z passes):
    lib.test_filter on a 512^3 box of noise + a delta function against the compiled reference."""
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    n = 512
    inputs = common.make_inputs(hii=n, dim=n, box_len=768.0)
    rng = np.random.default_rng(11)
    box = rng.normal(size=(n,) * 3).astype(np.float32)
    box[n // 2, n // 3, 5] += 100.0
    out = np.zeros((n,) * 3, np.float64)
    exp = np.zeros((n,) * 3, np.float64)
    gpu.state.init(inputs, broadcast_inputs=True)
    assert gpu.lib.test_filter(box.ctypes.data_as(C.POINTER(C.c_float)), 12.0, 0.0, 0.0, filter_flag,
                               out.ctypes.data_as(C.POINTER(C.c_double))) == 0
    ref.state.init(inputs, broadcast_inputs=True)
    assert ref.lib.test_filter(box.ctypes.data_as(C.POINTER(C.c_float)), 12.0, 0.0, 0.0, filter_flag,
                               exp.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert np.abs(out - exp).max() <= 5e-6 * np.abs(exp).max()


def test_device_gaussian_stream(gpu):
    """The reference's mt19937 polar-Gaussian stream generated on the GPU (csrc/gslrng.cu) against the
    sequential host generator: same values in the same order (the device's log() may differ from
    glibc's in the last bit), across two consecutive requests (carried-over words)."""
    lib = gpu.lib
    lib.b200_gsl_gaussian_stream.argtypes = [C.c_ulong, C.c_longlong, C.c_longlong, C.POINTER(C.c_double)]
    lib.b200_host_gaussian_stream.argtypes = [C.c_ulong, C.c_longlong, C.POINTER(C.c_double)]
    for seed, n1, n2 in ((12345, 3_000_000, 1_000_001), (7, 17, 5)):
        a = np.zeros(n1 + n2); b = np.zeros(n1 + n2)
        assert lib.b200_gsl_gaussian_stream(seed, n1, n2, a.ctypes.data_as(C.POINTER(C.c_double))) == 0
        assert lib.b200_host_gaussian_stream(seed, n1 + n2, b.ctypes.data_as(C.POINTER(C.c_double))) == 0
        assert np.abs(a - b).max() <= 1e-14 * max(1.0, np.abs(b).max()), np.abs(a - b).max()
        assert (a == b).mean() > 0.99


def test_smoothed_evolved_density_vs_reference(gpu):
    """SMOOTH_EVOLVED_DENSITY_FIELD=True (PerturbedField.c:221-227): gaussian smoothing of the evolved
    field in k space, velocities derived from the smoothed box."""
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    inputs = common.make_inputs(hii=64, dim=128, smooth_evolved=True)
    r_ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    r_pf = pkg.perturb_field(redshift=7.0, initial_conditions=r_ics, backend=ref)
    pf = pkg.perturb_field(redshift=7.0, initial_conditions=r_ics, backend=gpu)
    common.compare_struct(pf, r_pf, tols={"velocity_z": common.TOL_VELOCITY})
    plain = pkg.perturb_field(redshift=7.0, initial_conditions=pkg.compute_initial_conditions(
        inputs=common.make_inputs(hii=64, dim=128), backend=ref), backend=gpu)
    assert pf.density.std() < plain.density.std()   # the smoothing did something


def _tophat_profile(r, R):
    return np.where(r < R, 3.0 / (4 * np.pi * R**3), 0.0)


@pytest.mark.parametrize("R", [1.5, 5, 10, 20])
@pytest.mark.parametrize("filter_flag", [0, 1, 2, 3, 4])
def test_filter_known_answer(gpu, R, filter_flag):
    """Reference tests/test_filtering.py:111-236: a delta function through lib.test_filter keeps
    its sum to 1e-4 and (top-hat) follows the continuous-space profile."""
    inputs = common.make_inputs(hii=50, dim=100, box_len=100.0)
    gpu.state.init(inputs, broadcast_inputs=True)
    box = np.zeros((50, 50, 50), np.float32)
    box[25, 25, 25] = 1.0
    out = np.zeros((50, 50, 50), np.float64)
    R_param = {3: 20.0, 4: R / 2}.get(filter_flag, 0.0)
    st = gpu.lib.test_filter(box.ctypes.data_as(C.POINTER(C.c_float)), float(R), float(R_param), 0.0,
                             filter_flag, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert st == 0
    if filter_flag in (0, 1, 2, 4):
        assert abs(out.sum() - 1.0) < 1e-4
    ref = common.ref_backend()
    if ref is not None:
        ref.state.init(inputs, broadcast_inputs=True)
        exp = np.zeros_like(out)
        ref.lib.test_filter(box.ctypes.data_as(C.POINTER(C.c_float)), float(R), float(R_param), 0.0,
                            filter_flag, exp.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.abs(out - exp).max() <= 2e-6 * np.abs(exp).max() + 1e-9


@pytest.mark.parametrize("perturb", ["2LPT", "ZELDOVICH", "LINEAR"])
def test_perturb_roll_known_answer(gpu, perturb):
    """Reference tests/test_perturb.py:41-135: uniform IC velocities chosen to move the mass by
    exactly one low-res cell (+y for ZA, additionally -z for the 2LPT term) roll the density."""
    lo, hi, box_len, z = 4, 12, 8.0, 8.0
    inputs = common.make_inputs(hii=lo, dim=hi, box_len=box_len, perturb=perturb)
    gpu.state.init(inputs, broadcast_inputs=True)
    d_z, d_i = gpu.lib.dicke(z), gpu.lib.dicke(inputs.simulation_options.INITIAL_REDSHIFT)
    cell = box_len / lo
    ics = pkg.InitialConditions.new(inputs)
    ics.lowres_vy[...] = cell / (d_z - d_i)
    if perturb == "2LPT":
        ics.lowres_vz_2LPT[...] = cell / ((-3.0 / 7.0) * (d_z**2 - d_i**2))
    ics.lowres_density[0, 0, 0] = 1
    ics.lowres_density[lo // 2, lo // 2, lo // 2] = -1
    ics.hires_density[0, 0, 0] = 3**3
    ics.hires_density[hi // 2, hi // 2, hi // 2] = -(3**3)
    roll = {"LINEAR": (0, 0, 0), "ZELDOVICH": (0, 1, 0), "2LPT": (0, 1, -1)}[perturb]
    expected = np.roll(ics.lowres_density, roll, (0, 1, 2)) * (d_z if perturb == "LINEAR" else d_i)
    pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=gpu)
    np.testing.assert_allclose(pf.density, expected, atol=1e-3)


def test_bit_reproducible(gpu):
    """Two runs on the same inputs are bit-identical (fixed-point CIC, ordered reductions)."""
    inputs, ics, g_pf, _ = common.load_golden(common.GOLDEN_BASE)
    a = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=gpu)
    b = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=gpu)
    assert np.array_equal(a.density, b.density) and np.array_equal(a.velocity_z, b.velocity_z)
    ia = pkg.compute_ionization_field(perturbed_field=a, initial_conditions=ics, backend=gpu)
    ib = pkg.compute_ionization_field(perturbed_field=a, initial_conditions=ics, backend=gpu)
    for k in ia.arrays():
        assert np.array_equal(ia.arrays()[k], ib.arrays()[k]), k


def test_full_size_properties(gpu):
    """HII_DIM=256 (BASELINE config C2 size): filter linearity and sum conservation, mass
    conservation of the CIC deposit, monotone ionisation in the efficiency."""
    hii = 256
    inputs = common.make_inputs(hii=hii, dim=512, box_len=300.0)
    gpu.state.init(inputs, broadcast_inputs=True)
    rng = np.random.default_rng(7)
    a = rng.normal(size=(hii,) * 3).astype(np.float32)
    b = rng.normal(size=(hii,) * 3).astype(np.float32)

    def filt(x, R=7.0):
        out = np.zeros(x.shape, np.float64)
        assert gpu.lib.test_filter(x.ctypes.data_as(C.POINTER(C.c_float)), R, 0.0, 0.0, 0,
                                   out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        return out
    fa, fb, fab = filt(a), filt(b), filt((a + b).astype(np.float32))
    assert np.abs(fab - fa - fb).max() < 5e-6 * np.abs(fab).max() + 1e-6   # linearity
    assert abs(fa.sum() - a.astype(np.float64).sum()) < 1e-3 * hii**1.5      # DC mode is kept
    assert fa.std() < a.std()                                                # smoothing


# ---------------------------------------------------------------------------------------------
# Parity at the sizes the benchmark runs (BASELINE.json configs C2 and C3 / the bench default)
# ---------------------------------------------------------------------------------------------
def _ref_threads(inputs):
    import os
    return inputs.evolve_input_structs(N_THREADS=os.cpu_count() or 1)


def _perturb_ionize_vs_reference(gpu, ref, inputs, ics, z):
    """the same host ICs through both shared libraries; returns the comparison statistics"""
    pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=gpu)
    ics.inputs = _ref_threads(inputs)
    r_pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=ref)
    e_pf = common.compare_struct(pf, r_pf, tols={"velocity_z": common.TOL_VELOCITY})
    r_ib = pkg.compute_ionization_field(perturbed_field=r_pf, initial_conditions=ics, backend=ref)
    ics.inputs = inputs
    r_pf.inputs = inputs
    ib = pkg.compute_ionization_field(perturbed_field=r_pf, initial_conditions=ics, backend=gpu)
    stats = common.compare_ionized(ib, r_ib)
    # The flag of a cell is f_coll(delta_R) zeta > 1 with delta_R out of a float32 FFT: the two FFTs (this
    # library's Stockham passes, the oracle's MKL shim) agree to ~2e-7, so a cell whose value sits within that
    # rounding band of the threshold at some radius can flip -- as it does between two FFTW builds.  At these
    # sizes (1.7e7 / 1.3e8 cells x 32 / 40 radii) a handful of such cells exists; everything else is bit-exact.
    assert stats["mask_mismatch"] <= max(2, int(1e-6 * r_ib.neutral_fraction.size)), stats
    assert 0.05 < float(r_ib.neutral_fraction.mean()) < 0.95  # a partially ionised box: the ladder did work
    return e_pf, stats


def test_config_c2_vs_reference(gpu):
    """BASELINE.json config C2 exactly: perturb_field + ionize_box, z = 8, HII_DIM = 256, DIM = 768,
    BOX_LEN = 300 Mpc (32 filter radii), GPU-made ICs (exact GSL stream) fed to both libraries."""
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    inputs = common.make_inputs(hii=256, dim=768, box_len=300.0)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=gpu)
    e_pf, stats = _perturb_ionize_vs_reference(gpu, ref, inputs, ics, 8.0)
    print("C2 parity:", e_pf, stats)


def test_bench_workload_512_vs_reference(gpu, monkeypatch):
    """The workload bench.py times (BASELINE config C3's grid): HII_DIM = 512, DIM = 1536, BOX_LEN = 768,
    R_BUBBLE_MAX = 40 -> 40 filter radii, E-INTEGRAL, on the library's DEFAULT path (F = 3 grouped deposit
    behind the pipelined upload, tabulated window rows, single-precision flag sweep, transform run-ahead)
    against the compiled reference: ionised mask bit-exact, fields within TOL_FIELD."""
    import psutil
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    dim = 1536 if psutil.virtual_memory().available > 90 * 2**30 else 1024
    monkeypatch.setenv("B200_IC_RNG", "device")          # Philox field: the ICs are inputs here, not under test
    monkeypatch.setenv("B200_SKIP_SCRATCH_OUTPUTS", "1")  # no 3 x DIM^3 scratch boxes on the host
    inputs = common.make_inputs(hii=512, dim=dim, box_len=768.0, R_BUBBLE_MAX=40.0)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=gpu)
    e_pf, stats = _perturb_ionize_vs_reference(gpu, ref, inputs, ics, 8.0)
    print(f"512 / DIM={dim} parity:", e_pf, stats)


@pytest.mark.parametrize("hii,dim,source", [(64, 128, "E-INTEGRAL"), (128, 384, "CONST-ION-EFF")])
def test_device_entry_points_equal_host_entry(gpu, hii, dim, source):
    """b200_ComputePerturbedField_device / b200_ComputeIonizedBox_device (what bench.py's `value` leg
    calls, device pointers) are bit-identical to the reference-facing host-pointer entry points."""
    import torch
    from importlib import import_module
    _abi = import_module("21cmfast_b200._abi")
    inputs = common.make_inputs(hii=hii, dim=dim, source=source)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=gpu)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=gpu)
    ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=gpu)
    lib = gpu.lib
    dev = torch.device("cuda", 0)
    names_ic = ["hires_density", "lowres_density", "lowres_vx", "lowres_vy", "lowres_vz",
                "lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
    d_ic = {k: torch.from_numpy(getattr(ics, k)).to(dev) for k in names_ic}
    d_pf = {k: torch.zeros((hii,) * 3, dtype=torch.float32, device=dev) for k in ("density", "velocity_z")}
    d_ib = {k: torch.zeros((hii,) * 3, dtype=torch.float32, device=dev)
            for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion")}
    d_ib["neutral_fraction"].fill_(1.0)
    s_ic, s_pf, s_ib = _abi.InitialConditionsStruct(), _abi.PerturbedFieldStruct(), _abi.IonizedBoxStruct()
    for s, d in ((s_ic, d_ic), (s_pf, d_pf), (s_ib, d_ib)):
        for k, t in d.items():
            setattr(s, k, C.cast(t.data_ptr(), _abi.c_float_p))
    lib.b200_ComputePerturbedField_device.argtypes = [C.c_float, C.POINTER(_abi.InitialConditionsStruct),
                                                      C.POINTER(_abi.PerturbedFieldStruct)]
    lib.b200_ComputeIonizedBox_device.argtypes = [C.c_float, C.c_float, C.POINTER(_abi.PerturbedFieldStruct),
                                                  C.POINTER(_abi.IonizedBoxStruct)]
    torch.cuda.synchronize()
    gpu.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True)
    assert lib.b200_ComputePerturbedField_device(C.c_float(8.0), C.byref(s_ic), C.byref(s_pf)) == 0
    assert lib.b200_ComputeIonizedBox_device(C.c_float(8.0), C.c_float(-1.0), C.byref(s_pf), C.byref(s_ib)) == 0
    for k, t in d_pf.items():
        assert np.array_equal(t.cpu().numpy(), getattr(pf, k)), k
    for k, t in d_ib.items():
        assert np.array_equal(t.cpu().numpy(), getattr(ib, k).reshape(t.shape)), k
    assert s_ib.mean_f_coll == ib.mean_f_coll
