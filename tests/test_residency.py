"""Residency of the boxes handed from one of the reference's entry points to the next (SURVEY.md section 8f row 1;
drivers/coeval.py:835-853: PerturbedField -> IonizedBox -> BrightnessTemp): with b200_residency(1) the consumers
use the device copies the producers left behind instead of uploading the host arrays again -- same results,
fewer bytes over PCIe; off by default, and dropped when switched off."""
import ctypes as C

import numpy as np
import pytest

import common

pkg = common.pkg


def _stats(be):
    a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_double()
    be.lib.b200_last_call_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    return b.value, c.value


def _chain(be, ics, z):
    h2d = {}
    pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=be)
    ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be)
    h2d["ionize"] = _stats(be)[0]
    bt = pkg.brightness_temperature(ionized_box=ib, perturbed_field=pf, backend=be)
    h2d["tb"] = _stats(be)[0]
    return pf, ib, bt, h2d


def _check(be):
    be.lib.b200_residency.argtypes, be.lib.b200_residency.restype = [C.c_int], None
    inputs = common.make_inputs(hii=32, dim=64, source="E-INTEGRAL")
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    n_bytes = 4 * 32**3
    pf0, ib0, bt0, plain = _chain(be, ics, 8.0)
    be.lib.b200_residency(1)
    try:
        pf1, ib1, bt1, res = _chain(be, ics, 8.0)
        # a second redshift while the first one's copies are still cached: keys are the host arrays, not the call order
        pf2, ib2, bt2, _ = _chain(be, ics, 9.0)
        bt1_again = pkg.brightness_temperature(ionized_box=ib1, perturbed_field=pf1, backend=be)
    finally:
        be.lib.b200_residency(0)
    for a, b in ((pf0, pf1), (ib0, ib1), (bt0, bt1), (bt0, bt1_again)):
        for k, v in a.arrays().items():
            assert np.array_equal(v, b.arrays()[k]), k
    assert not np.array_equal(ib2.neutral_fraction, ib1.neutral_fraction)
    # the perturbed density is not uploaded again by ComputeIonizedBox, nor density + neutral fraction by T_b
    assert plain["ionize"] - res["ionize"] == n_bytes, (plain, res)
    assert plain["tb"] - res["tb"] == 2 * n_bytes, (plain, res)
    # switched off again: every call reads the caller's arrays, e.g. an edited one
    pf1.density[...] = pf1.density * 0.5
    ib3 = pkg.compute_ionization_field(perturbed_field=pf1, initial_conditions=ics, backend=be)
    assert _stats(be)[0] == plain["ionize"]
    assert not np.array_equal(ib3.neutral_fraction, ib1.neutral_fraction)


def test_residency_between_entry_points_emulated():
    be = common.emu_backend()
    if be is None:
        pytest.skip("tests/_emu not built")
    _check(be)


@pytest.mark.gpu
def test_residency_between_entry_points_gpu():
    _check(common.gpu_backend())
