"""Seed parity of the initial conditions with N_THREADS > 1: the
reference gives each OpenMP thread its own GSL generator -- mt19937, gfsr4, cmrg, mrg, taus2 by
thread index (rng.c:57-83), seeded by a choose + shuffle of the user seed that depends on N_THREADS
(rng.c:31-54) -- and a contiguous range of x planes (static schedule, InitialConditions.c:103-134).
The thread counts below reach every generator type, a second mt19937 (thread 5) and uneven ranges."""
import numpy as np
import pytest

import common

pkg = common.pkg

# (N_THREADS, HII_DIM, DIM)
CASES = [(2, 16, 32), (3, 16, 32), (5, 12, 36), (6, 16, 32), (7, 10, 30)]


def _run(be, ref, nt, hii, dim):
    inputs = common.make_inputs(hii=hii, dim=dim, seed=4321, n_threads=nt)
    got = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    want = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    common.compare_struct(got, want)
    # a different thread count is a different field (the seeds and the plane ranges both change)
    other = pkg.compute_initial_conditions(inputs=common.make_inputs(hii=hii, dim=dim, seed=4321, n_threads=1), backend=be)
    assert np.abs(other.hires_density - got.hires_density).max() > 0.1 * np.abs(got.hires_density).max()


@pytest.mark.parametrize("nt,hii,dim", CASES)
def test_ic_seed_parity_with_threads_emulated(nt, hii, dim):
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    _run(emu, ref, nt, hii, dim)


@pytest.mark.gpu
@pytest.mark.parametrize("nt,hii,dim", [(3, 16, 32), (7, 10, 30)])
def test_ic_seed_parity_with_threads_gpu(nt, hii, dim):
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    _run(common.gpu_backend(), ref, nt, hii, dim)


def test_static_schedule_ranges_cover_the_grid():
    """libgomp's static schedule: contiguous, ordered, the first n % T threads one plane longer."""
    import ctypes as C
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    fn = emu.lib.b200_omp_static_range
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    fn.restype = None
    for n, T in ((30, 7), (32, 3), (5, 8), (64, 64), (1536, 16)):
        pos, sizes = 0, []
        for t in range(T):
            b, e = C.c_int(), C.c_int()
            fn(n, T, t, C.byref(b), C.byref(e))
            assert b.value == pos and e.value >= b.value
            pos = e.value
            sizes.append(e.value - b.value)
        assert pos == n and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
