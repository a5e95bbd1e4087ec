"""Pin the oracle's third-party restatements (oracle/shims) against independent implementations:
MT19937 vs numpy's legacy seeding, the polar Gaussian vs a numpy re-derivation from the same
uniform stream, QAG-61 vs scipy's QUADPACK, the natural cubic spline vs scipy, and both FFT
back-ends vs numpy.  CPU only."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "oracle" / "_ref" / "liboracle_shims.so"


@pytest.fixture(scope="module")
def shim():
    if not SHIM.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "shims"], check=True, capture_output=True)
    lib = C.CDLL(str(SHIM))
    lib.gsl_rng_alloc.restype = C.c_void_p
    lib.gsl_rng_alloc.argtypes = [C.c_void_p]
    lib.gsl_rng_set.argtypes = [C.c_void_p, C.c_ulong]
    lib.gsl_rng_get.argtypes = [C.c_void_p]
    lib.gsl_rng_get.restype = C.c_ulong
    lib.gsl_rng_uniform.argtypes = [C.c_void_p]
    lib.gsl_rng_uniform.restype = C.c_double
    lib.gsl_ran_ugaussian.argtypes = [C.c_void_p]
    lib.gsl_ran_ugaussian.restype = C.c_double
    return lib


def _rng(lib, name, seed):
    T = C.c_void_p.in_dll(lib, name)
    r = lib.gsl_rng_alloc(T)
    lib.gsl_rng_set(r, seed)
    return r


@pytest.mark.parametrize("seed", [1, 12345, 4357, 2**31 + 7])
def test_mt19937_matches_numpy_legacy_seeding(shim, seed):
    r = _rng(shim, "gsl_rng_mt19937", seed)
    got = np.array([shim.gsl_rng_get(r) for _ in range(2000)], dtype=np.uint64)
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    exp = bg.random_raw(2000)
    assert np.array_equal(got, exp)


def test_mt19937_seed_zero_is_4357(shim):
    a, b = _rng(shim, "gsl_rng_mt19937", 0), _rng(shim, "gsl_rng_mt19937", 4357)
    assert [shim.gsl_rng_get(a) for _ in range(10)] == [shim.gsl_rng_get(b) for _ in range(10)]


def test_polar_gaussian_follows_gsl_definition(shim):
    """gsl_ran_gaussian: x,y = -1 + 2 U_pos; accept r2 in (0,1]; return y sqrt(-2 ln r2 / r2)."""
    r = _rng(shim, "gsl_rng_mt19937", 99)
    got = np.array([shim.gsl_ran_ugaussian(r) for _ in range(500)])
    bg = np.random.MT19937()
    bg._legacy_seeding(99)
    raw = iter(bg.random_raw(5000) / 4294967296.0)

    def upos():
        while True:
            u = next(raw)
            if u != 0:
                return u
    exp = []
    for _ in range(500):
        while True:
            x, y = -1 + 2 * upos(), -1 + 2 * upos()
            r2 = x * x + y * y
            if 0 < r2 <= 1:
                break
        exp.append(y * np.sqrt(-2 * np.log(r2) / r2))
    np.testing.assert_allclose(got, exp, rtol=1e-15, atol=0)
    assert abs(got.mean()) < 0.15 and abs(got.std() - 1) < 0.1


def test_qag61_matches_quadpack(shim):
    from scipy import integrate

    class F(C.Structure):
        _fields_ = [("function", C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)), ("params", C.c_void_p)]
    shim.gsl_integration_workspace_alloc.restype = C.c_void_p
    shim.gsl_integration_workspace_alloc.argtypes = [C.c_size_t]
    shim.gsl_integration_qag.argtypes = [C.POINTER(F), C.c_double, C.c_double, C.c_double, C.c_double, C.c_size_t,
                                         C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    w = shim.gsl_integration_workspace_alloc(1000)
    cases = [(lambda x: np.exp(-x * x) * np.cos(3 * x), -2.0, 5.0), (lambda x: 1 / (1e-3 + x * x), -1.0, 1.0),
             (lambda x: np.sqrt(x) * np.log(x + 1e-9), 0.0, 3.0), (lambda x: x**7 - 3 * x**2, -1.0, 2.0)]
    for fn, a, b in cases:
        cb = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)(lambda x, p, fn=fn: float(fn(x)))
        f = F(cb, None)
        res, err = C.c_double(), C.c_double()
        st = shim.gsl_integration_qag(C.byref(f), a, b, 0.0, 1e-10, 1000, 6, w, C.byref(res), C.byref(err))
        assert st == 0
        exp, _ = integrate.quad(fn, a, b, epsabs=0, epsrel=1e-12, limit=500)
        assert abs(res.value - exp) <= 2e-10 * max(1.0, abs(exp))
    # a 61-point rule integrates a degree-91 polynomial exactly on one panel
    cb = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)(lambda x, p: x**90)
    f = F(cb, None)
    res, err = C.c_double(), C.c_double()
    shim.gsl_integration_qag(C.byref(f), -1.0, 1.0, 0.0, 1e-3, 1, 6, w, C.byref(res), C.byref(err))
    assert abs(res.value - 2 / 91) < 1e-14


def test_natural_cspline_matches_scipy(shim):
    from scipy.interpolate import CubicSpline
    shim.gsl_spline_alloc.restype = C.c_void_p
    shim.gsl_spline_alloc.argtypes = [C.c_void_p, C.c_size_t]
    shim.gsl_spline_init.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t]
    shim.gsl_spline_eval.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    shim.gsl_spline_eval.restype = C.c_double
    x = np.sort(np.random.default_rng(0).uniform(0, 10, 40))
    y = np.sin(x) + 0.1 * x
    sp = shim.gsl_spline_alloc(C.c_void_p.in_dll(shim, "gsl_interp_cspline"), len(x))
    shim.gsl_spline_init(sp, x.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double)), len(x))
    ref = CubicSpline(x, y, bc_type="natural")
    xs = np.linspace(x[0], x[-1], 300)
    got = np.array([shim.gsl_spline_eval(sp, float(v), None) for v in xs])
    np.testing.assert_allclose(got, ref(xs), rtol=1e-11, atol=1e-12)


def _gamma_inc_cases():
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(-11, 1, 160), [-4.75, -9.95, -3.0, -1.0, 0.0, 0.5, -0.5, -10.0]])
    x = np.concatenate([10 ** rng.uniform(-6, np.log10(300), 160), [80.0, 300.0, 300.0, 0.2, 0.3, 1e-6, 0.25, 150.0]])
    return a, x


def _check_gamma_inc(fn):
    """upper incomplete gamma for real a <= 1 against mpmath (the GAMMA-APPROX integrals call it with
    a = 1/2 + beta down to about -10 and x up to a few hundred, hmf.c:728-760)"""
    import mpmath
    mpmath.mp.dps = 40
    a, x = _gamma_inc_cases()
    worst = 0.0
    for ai, xi in zip(a, x):
        want = float(mpmath.gammainc(mpmath.mpf(float(ai)), mpmath.mpf(float(xi))))
        got = fn(float(ai), float(xi))
        if want == 0.0 or not np.isfinite(want):
            continue
        worst = max(worst, abs(got - want) / abs(want))
    assert worst < 2e-11, worst


def test_gamma_inc_shim_matches_mpmath(shim):
    shim.gsl_sf_gamma_inc.argtypes = [C.c_double, C.c_double]
    shim.gsl_sf_gamma_inc.restype = C.c_double
    _check_gamma_inc(shim.gsl_sf_gamma_inc)


def test_gamma_inc_product_matches_mpmath():
    """the shipped library's own restatement (host code, no GPU needed) against the same oracle"""
    lib_path = ROOT / "21cmfast_b200" / "csrc" / "lib21cmfast_b200.so"
    if not lib_path.exists():
        pytest.skip("product library not built")
    lib = C.CDLL(str(lib_path))
    lib.b200_upper_gamma.argtypes = [C.c_double, C.c_double]
    lib.b200_upper_gamma.restype = C.c_double
    _check_gamma_inc(lib.b200_upper_gamma)


@pytest.mark.parametrize("backend", ["own", "mkl"])
@pytest.mark.parametrize("shape", [(16, 16, 16), (12, 10, 14), (35, 35, 35), (8, 8, 30)])
def test_fftw_shim_matches_numpy(backend, shape):
    """r2c / c2r of the FFTW shim (both back-ends) vs numpy, in a subprocess because the back-end
    is chosen once per process."""
    code = f"""
import ctypes as C, numpy as np, os, sys
os.environ['ORACLE_FFT'] = '{backend}'
if '{backend}' == 'mkl':
    import torch
    os.environ['ORACLE_TORCH_LIB'] = os.path.join(os.path.dirname(torch.__file__), 'lib', 'libtorch_cpu.so')
lib = C.CDLL('{SHIM}')
lib.fftwf_plan_dft_r2c_3d.restype = C.c_void_p; lib.fftwf_plan_dft_c2r_3d.restype = C.c_void_p
lib.fftwf_plan_dft_r2c_3d.argtypes = [C.c_int]*3 + [C.c_void_p, C.c_void_p, C.c_uint]
lib.fftwf_plan_dft_c2r_3d.argtypes = [C.c_int]*3 + [C.c_void_p, C.c_void_p, C.c_uint]
lib.fftwf_execute.argtypes = [C.c_void_p]
n0, n1, n2 = {shape}
nc = n2 // 2 + 1
rng = np.random.default_rng(3)
x = rng.normal(size=(n0, n1, n2)).astype(np.float32)
buf = np.zeros((n0, n1, 2 * nc), np.float32); buf[:, :, :n2] = x
p = lib.fftwf_plan_dft_r2c_3d(n0, n1, n2, buf.ctypes.data, buf.ctypes.data, 64); lib.fftwf_execute(p)
got = buf.view(np.complex64).reshape(n0, n1, nc)
exp = np.fft.rfftn(x.astype(np.float64))
assert np.abs(got - exp).max() < 2e-5 * np.abs(exp).max(), 'r2c'
p = lib.fftwf_plan_dft_c2r_3d(n0, n1, n2, buf.ctypes.data, buf.ctypes.data, 64); lib.fftwf_execute(p)
back = buf[:, :, :n2] / (n0 * n1 * n2)
assert np.abs(back - x).max() < 2e-5, 'c2r'
print('mkl' if lib.oracle_fft_backend_is_mkl() else 'own')
"""
    if not SHIM.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "shims"], check=True, capture_output=True)
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    if backend == "own":
        assert r.stdout.strip() == "own"


@pytest.mark.parametrize("backend", ["own", "mkl"])
def test_fftw_shim_c2r_of_non_hermitian_input_follows_fftw(backend):
    """The reference hands c2r planes that are not Hermitian (i k at the Nyquist index: InitialConditions.c
    velocity modes, PerturbedField.c:344-345).  FFTW's answer is fixed by its algorithm: complex transforms over
    axes 0 and 1 of the stored half, then the real transform along axis 2 (imaginary parts of the self-conjugate
    bins dropped) -- numpy's ifft2 + irfft below.  MKL alone differs on small non-cubic shapes (0.1 relative on
    12x12x15); the shim probes such shapes once and runs them on its own FFT."""
    code = f"""
import ctypes as C, numpy as np, os
os.environ['ORACLE_FFT'] = '{backend}'
if '{backend}' == 'mkl':
    import torch
    os.environ['ORACLE_TORCH_LIB'] = os.path.join(os.path.dirname(torch.__file__), 'lib', 'libtorch_cpu.so')
lib = C.CDLL('{SHIM}')
lib.fftwf_plan_dft_c2r_3d.restype = C.c_void_p
lib.fftwf_plan_dft_c2r_3d.argtypes = [C.c_int]*3 + [C.c_void_p, C.c_void_p, C.c_uint]
lib.fftwf_execute.argtypes = [C.c_void_p]
for n0, n1, n2 in [(12, 12, 15), (10, 10, 12), (8, 8, 30), (8, 8, 7), (16, 16, 16), (12, 12, 18), (24, 24, 36), (50, 50, 50)]:
    nc = n2 // 2 + 1
    rng = np.random.default_rng(3)
    g = (rng.normal(size=(n0, n1, nc)) + 1j * rng.normal(size=(n0, n1, nc))).astype(np.complex64)
    buf = np.zeros((n0, n1, 2 * nc), np.float32)
    buf.view(np.complex64).reshape(n0, n1, nc)[...] = g
    p = lib.fftwf_plan_dft_c2r_3d(n0, n1, n2, buf.ctypes.data, buf.ctypes.data, 64)
    lib.fftwf_execute(p)
    exp = np.fft.irfft(np.fft.ifft2(g.astype(np.complex128), axes=(0, 1)), n=n2, axis=2) * (n0 * n1 * n2)
    err = np.abs(buf[:, :, :n2] - exp).max() / np.abs(exp).max()
    assert err < 2e-6, ((n0, n1, n2), err)
"""
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
