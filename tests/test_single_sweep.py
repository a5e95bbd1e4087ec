"""The single-sweep radius ladder (ionize.cu, SpecState): the sum sweep flags cells against a bracket of the
predicted mean fix and queues the cells inside it; spec_resolve_kernel decides those once the grid sum is
known.  It replaces the second walk of find_ionised_regions (IonisationBox.c:1008-1201) over the filtered
grid, so every output must be BIT-identical to the two-sweep ladder -- also when a queue segment overflows
(that radius falls back to the full flag sweep) and when the prediction leaves its bracket (the ladder is
re-run without speculation)."""
import os

import numpy as np
import pytest

import common

pkg = common.pkg


def _ladder(be, pf, ics, **env):
    keys = ("B200_SPEC", "B200_SPEC_QCAP", "B200_SPEC_EPS", "B200_FUSED_LAST")
    old = {k: os.environ.pop(k, None) for k in keys}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be)
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if old[k] is not None:
                os.environ[k] = old[k]


def _identical(a, b):
    for k, v in a.arrays().items():
        assert np.array_equal(v, b.arrays()[k]), k
    assert a.mean_f_coll == b.mean_f_coll


def _check(be, ics_be, source, hii, z):
    inputs = common.make_inputs(hii=hii, dim=2 * hii, source=source)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=ics_be)
    pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=ics_be)
    two_sweeps = _ladder(be, pf, ics, B200_SPEC=0)
    assert 0.05 < two_sweeps.global_xH < 0.98  # a box that is being ionised, not a trivial one
    _identical(_ladder(be, pf, ics), two_sweeps)
    _identical(_ladder(be, pf, ics, B200_SPEC_QCAP=2), two_sweeps)      # every segment overflows
    _identical(_ladder(be, pf, ics, B200_SPEC_EPS=1e-7), two_sweeps)    # the prediction must fail: re-run
    # the last radius and the finalisation as one pass (default) or as two kernels
    _identical(_ladder(be, pf, ics, B200_SPEC=0, B200_FUSED_LAST=0), two_sweeps)
    return two_sweeps


@pytest.mark.parametrize("source,z", [("E-INTEGRAL", 8.0), ("CONST-ION-EFF", 8.0)])
def test_single_sweep_ladder_is_bit_identical_emulated(source, z):
    be = common.emu_backend()
    if be is None:
        pytest.skip("tests/_emu not built")
    _check(be, be, source, 32, z)


@pytest.mark.gpu
@pytest.mark.parametrize("source,hii,z", [("E-INTEGRAL", 128, 8.0), ("CONST-ION-EFF", 64, 8.0), ("E-INTEGRAL", 64, 6.5)])
def test_single_sweep_ladder_is_bit_identical_gpu(source, hii, z):
    be = common.gpu_backend()
    _check(be, be, source, hii, z)
