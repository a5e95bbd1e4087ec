"""POWER_SPECTRUM = 'CLASS' and V_CB_MODEL = 'FLUCTS' (SURVEY.md section 8 row a7): the tabulated transfer
functions (transfer_function_CLASS, cosmology.c:130-213; power_in_k / power_in_vcb, :278-332) and
compute_relative_velocities (InitialConditions.c:141-238) against the compiled reference.  classy is not part of
this build, so both sides get the same synthetic tables: an Eisenstein & Hu-like density transfer function in the
CLASS convention (T = delta / zeta ~ k^2 T_EH) and a smooth relative-velocity transfer function."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

import common

pkg = common.pkg
CosmoTables = __import__("importlib").import_module("21cmfast_b200.inputs").CosmoTables


def _tables(kmax=40.0, n=160):
    k = np.concatenate(([0.0], np.geomspace(1e-5, kmax, n - 1)))
    q = k / 0.14
    bbks = np.ones_like(k)
    bbks[1:] = np.log(1 + 2.34 * q[1:]) / (2.34 * q[1:]) * (1 + 3.89 * q[1:] + (16.1 * q[1:]) ** 2 + (5.46 * q[1:]) ** 3
                                                          + (6.71 * q[1:]) ** 4) ** -0.25
    t_m = 3.0e3 * k * k * bbks                  # grows like k^2 at small k, bends over at the horizon scale
    t_v = 2.0e-2 * k / (1.0 + (k / 0.05) ** 2) * (1 + 0.3 * np.sin(k / 0.03) * np.exp(-k / 0.3))  # peaks near the BAO scale
    t_v[0] = 0.0
    return (k, t_m), (k, np.abs(t_v) + 1e-12 * k)


def _inputs(hii=24, dim=48, flucts=True, sigma8=True):
    base = common.make_inputs(hii=hii, dim=dim, seed=21, source="E-INTEGRAL")
    td, tv = _tables()
    ct = CosmoTables(ps_norm=base.cosmo_params.SIGMA_8 if sigma8 else 2.1e-9, USE_SIGMA_8=sigma8,
                     transfer_density=td, transfer_vcb=tv if flucts else None)
    mo = dataclasses.replace(base.matter_options, V_CB_MODEL="FLUCTS" if flucts else "NONE", POWER_SPECTRUM="CLASS")
    return dataclasses.replace(base, matter_options=mo, class_tables=ct)


def _scalars(be, inputs):
    be.state.init(inputs, broadcast_inputs=True, ps=True)
    out = {}
    for name in ("power_in_k", "power_in_vcb", "sigma_z0"):
        fn = getattr(be.lib, name)
        fn.restype, fn.argtypes = C.c_double, [C.c_double]
    ks = [1e-4, 3e-3, 0.05, 0.7, 12.0, 39.9, 55.0, 400.0]   # inside the table, at its end, beyond it
    out["pk"] = np.array([be.lib.power_in_k(k) for k in ks])
    if inputs.matter_options.V_CB_MODEL == "FLUCTS":
        out["pv"] = np.array([be.lib.power_in_vcb(k) for k in ks])
    out["sigma"] = np.array([be.lib.sigma_z0(m) for m in (1e8, 1e11, 1e14)])
    return out


def _check(be, flucts, sigma8):
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    inputs = _inputs(flucts=flucts, sigma8=sigma8)
    got, want = _scalars(be, inputs), _scalars(ref, inputs)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, err_msg=k)
    assert np.all(want["pk"] > 0)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    r_ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    errs = common.compare_struct(ics, r_ics, tol=common.TOL_FIELD)
    if flucts:
        assert float(r_ics.lowres_vcb.min()) >= 0 and float(r_ics.lowres_vcb.max()) > 0
        assert "lowres_vcb" in errs
    return errs


@pytest.mark.parametrize("flucts,sigma8", [(True, True), (False, True), (True, False)])
def test_class_tables_and_relative_velocities_emulated(flucts, sigma8):
    be = common.emu_backend()
    if be is None:
        pytest.skip("tests/_emu not built")
    print(_check(be, flucts, sigma8))


@pytest.mark.gpu
@pytest.mark.parametrize("flucts,sigma8", [(True, True), (False, True), (True, False)])
def test_class_tables_and_relative_velocities_gpu(flucts, sigma8):
    print(_check(common.gpu_backend(), flucts, sigma8))


def test_class_tables_after_another_binding_left_its_cosmo_tables_behind():
    """The C side copies the cosmo tables once per Free_cosmo_tables_global (InputParameters.c:9-53).  A second binding
    of the same shared object (cffi route B, another Backend) that broadcast analytic-spectrum inputs and never freed
    them must not leave a later CLASS run without its tables (this crashed the GPU tier once: the scalar entry points
    now also return NaN instead of dereferencing missing tables)."""
    path = common.ROOT / "tests" / "_emu" / "libb200_emu.so"
    if not path.exists() or common.ref_backend() is None:
        pytest.skip("tests/_emu / oracle/_ref not built")
    first = pkg.Backend(path)
    first.state.init(common.make_inputs(hii=16, dim=32), broadcast_inputs=True, ps=True)   # EH tables stay allocated
    second = pkg.Backend(path)
    inputs = _inputs(flucts=True, sigma8=True)
    got, want = _scalars(second, inputs), _scalars(common.ref_backend(), inputs)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, err_msg=k)
