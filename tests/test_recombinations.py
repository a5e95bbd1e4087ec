"""Recombinations, x_e filtering and previous-snapshot evolution of the IonizeBox path (SURVEY.md
section 8f, row 2) against the compiled reference (oracle/_ref): USE_TS_FLUCT with a caller-supplied
TsBox (the spin-temperature calculation itself is out of scope: both sides get the same synthetic
box), RECOMB_MODEL = homogeneous / inhomogeneous, with the
N_rec grid filtered per radius (CELL_RECOMB = False) or taken per cell, chained over three snapshots
so that z_reion, Gamma12, the mean free path and the cumulative recombinations of one snapshot feed
the next.  Also the rate table itself (init_MHR / splined_recombination_rate) as a known-answer test.
CPU tier: host-emulated kernels; GPU tier: the CUDA library."""
import dataclasses

import numpy as np
import pytest

import common

pkg = common.pkg

REDSHIFTS = (9.0, 8.0, 7.0)
CASES = {
    "inhomogeneous_filtered": dict(model="inhomogeneous", cell=False, source="E-INTEGRAL"),
    "inhomogeneous_cell": dict(model="inhomogeneous", cell=True, source="E-INTEGRAL"),
    "homogeneous_cell": dict(model="homogeneous", cell=True, source="E-INTEGRAL"),
    "inhomogeneous_filtered_const_zeta": dict(model="inhomogeneous", cell=False, source="CONST-ION-EFF"),
    "ts_fluct": dict(model="none", cell=False, source="E-INTEGRAL", ts=True),
    "ts_fluct_inhomogeneous_filtered": dict(model="inhomogeneous", cell=False, source="E-INTEGRAL", ts=True),
}
GOLDEN_CHAIN_CASES = ("inhomogeneous_filtered", "homogeneous_cell", "ts_fluct_inhomogeneous_filtered")
GOLDEN_FIELDS = ("neutral_fraction", "z_reion", "ionisation_rate_G12", "mean_free_path", "cumulative_recombinations",
                 "kinetic_temperature")
RATE_Z = np.array([-0.3, 0.0, 0.09, 0.11, 2.0, 5.37, 6.5, 8.0, 11.9, 13.0, 25.0, 59.8, 80.0])
RATE_GAMMA = np.exp(np.array([-12.0, -10.0, -9.95, -7.3, -2.0, -0.05, 0.0, 0.5, 3.1, 14.85, 14.9, 20.0]))
# the homogeneous model's one number comes from a float box sum in the reference (IonisationBox.c:1595-1607)
TOL_GLOBAL_NREC = 1e-4


def _inputs(model, cell, source, hii=32, ts=False):
    inp = common.make_inputs(hii=hii, dim=2 * hii, seed=77, source=source)
    ao = dataclasses.replace(inp.astro_options, RECOMB_MODEL=model, CELL_RECOMB=cell, USE_TS_FLUCT=ts)
    return dataclasses.replace(inp, astro_options=ao)


def _synthetic_ts(inputs, pf):
    """A TsBox with the right magnitudes and some structure: x_e of a few per cent following the density
    (with excursions beyond [0, 1] so that the clips act), adiabatic-like neutral-gas temperature, and a
    spin temperature between it and the CMB."""
    ts = pkg.TsBox.new(inputs, pf.redshift)
    rng = np.random.default_rng(int(pf.redshift * 100))
    d = pf.density.astype(np.float64)
    xe = 0.03 * (1 + d) + 0.02 * rng.standard_normal(d.shape)
    xe[rng.random(d.shape) < 1e-3] = 1.2
    ts.xray_ionised_fraction[...] = xe
    tk = 40.0 * np.cbrt(np.clip(1 + d, 1e-3, None)) ** 2 * (1 + 0.1 * rng.random(d.shape))
    ts.kinetic_temp_neutral[...] = tk
    ts.spin_temperature[...] = 0.5 * (tk + 2.7255 * (1 + pf.redshift)) + 1.0
    return ts


def _chain(be, inputs, ics, pfs):
    prev_ib, prev_pf = pkg.IonizedBox.initial(inputs), pkg.PerturbedField.initial(inputs)
    out = []
    for pf in pfs:
        ts = _synthetic_ts(inputs, pf) if inputs.astro_options.USE_TS_FLUCT else None
        ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, previous_ionized_box=prev_ib,
                                          previous_perturbed_field=prev_pf, spin_temp=ts, backend=be)
        out.append(ib)
        prev_ib, prev_pf = ib, pf
    return out


def _run_case(be, ref, name):
    kw = CASES[name]
    inputs = _inputs(**kw)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    pfs = [pkg.perturb_field(redshift=z, initial_conditions=ics, backend=ref) for z in REDSHIFTS]
    got, want = _chain(be, inputs, ics, pfs), _chain(ref, inputs, ics, pfs)
    for z, a, b in zip(REDSHIFTS, got, want):
        stats = common.compare_ionized(a, b)
        assert stats["mask_mismatch"] == 0, (name, z, stats)
        # the first crossing is recorded at the same radius: the mean free path is one of the ladder's radii, bit for bit
        assert np.array_equal(a.mean_free_path, b.mean_free_path), (name, z)
        assert np.array_equal(a.z_reion, b.z_reion), (name, z)
        for k, tol in (("ionisation_rate_G12", common.TOL_FIELD),
                       ("cumulative_recombinations", TOL_GLOBAL_NREC if kw["model"] == "homogeneous" else common.TOL_FIELD)):
            if kw["model"] == "none":
                break
            u, v = getattr(a, k), getattr(b, k)
            assert u.shape == v.shape and np.isfinite(u).all(), (name, z, k)
            err = np.abs(u - v).max() / max(np.abs(v).max(), 1e-30)
            assert err <= tol, (name, z, k, err)
    # the chain really evolved: reionisation redshifts of earlier snapshots survive, recombinations accumulate
    last = got[-1]
    assert set(np.unique(last.z_reion)) >= {-1.0, 9.0, 8.0, 7.0}
    if kw["model"] != "none":
        assert float(got[2].cumulative_recombinations.mean()) > float(got[1].cumulative_recombinations.mean()) > 0
        assert float(last.ionisation_rate_G12.max()) > 0 and float(last.mean_free_path.max()) > 0


@pytest.mark.parametrize("name", list(CASES))
def test_recombinations_emulated_vs_reference(name):
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    _run_case(emu, ref, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_recombinations_gpu_vs_reference(name):
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    _run_case(common.gpu_backend(), ref, name)


def _rate_table_case(be, ref):
    inputs = _inputs("inhomogeneous", False, "E-INTEGRAL", hii=16)
    zs, gammas = RATE_Z, RATE_GAMMA
    out = []
    for b in (be, ref):
        b.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True, recomb=True)
        out.append(np.array([[b.lib.splined_recombination_rate(float(z), float(g)) for g in gammas] for z in zs]))
    got, want = out
    assert np.isfinite(want).all() and (want[:, 0] == 0).all() and (want[:, 3:] > 0).all()
    assert np.allclose(got, want, rtol=1e-10, atol=0), np.abs(got / np.where(want == 0, 1, want) - 1).max()


def test_recombination_rate_table_emulated_vs_reference():
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    _rate_table_case(emu, ref)


@pytest.mark.gpu
def test_recombination_rate_table_gpu_vs_reference():
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    _rate_table_case(common.gpu_backend(), ref)


def test_recombinations_need_previous_boxes():
    """single_field.py:779-787: with evolution, both previous boxes are required below Z_HEAT_MAX."""
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    inputs = _inputs("inhomogeneous", False, "E-INTEGRAL", hii=16)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
    with pytest.raises(ValueError):
        pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=emu)
    with pytest.raises(ValueError):
        pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics,
                                     previous_ionized_box=pkg.IonizedBox.initial(inputs), backend=emu)


def _ts_neutral_and_brightness_case(be, ref):
    """USE_TS_FLUCT side branches: the fully neutral early exit (x_HI = 1 - x_e, T_k from the TsBox;
    IonisationBox.c:531-549) and the brightness temperature with a finite spin temperature and the
    21-cm optical depth (BrightnessTemperatureBox.c:73-82)."""
    inputs = _inputs("none", False, "E-INTEGRAL", hii=24, ts=True)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
    for z in (25.0, 8.0):
        pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=ref)
        ts = _synthetic_ts(inputs, pf)
        kw = dict(perturbed_field=pf, initial_conditions=ics, previous_ionized_box=pkg.IonizedBox.initial(inputs),
                  previous_perturbed_field=pkg.PerturbedField.initial(inputs), spin_temp=ts)
        a, b = pkg.compute_ionization_field(backend=be, **kw), pkg.compute_ionization_field(backend=ref, **kw)
        assert common.compare_ionized(a, b)["mask_mismatch"] == 0
        if z == 25.0:  # nothing ionised: the box is the TsBox's own x_e and temperature, bit for bit
            assert np.array_equal(a.neutral_fraction, b.neutral_fraction)
            assert np.array_equal(a.kinetic_temperature, ts.kinetic_temp_neutral)
        ta = pkg.brightness_temperature(ionized_box=b, perturbed_field=pf, spin_temp=ts, backend=be)
        tb = pkg.brightness_temperature(ionized_box=b, perturbed_field=pf, spin_temp=ts, backend=ref)
        for k in ("brightness_temp", "tau_21"):
            u, v = getattr(ta, k), getattr(tb, k)
            assert np.abs(v).max() > 0 and np.abs(u - v).max() <= 2e-6 * np.abs(v).max(), (z, k)
    with pytest.raises(ValueError):
        pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics,
                                     previous_ionized_box=pkg.IonizedBox.initial(inputs),
                                     previous_perturbed_field=pkg.PerturbedField.initial(inputs), backend=be)


def test_ts_fluct_neutral_box_and_brightness_emulated_vs_reference():
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    _ts_neutral_and_brightness_case(emu, ref)


@pytest.mark.gpu
def test_ts_fluct_neutral_box_and_brightness_gpu_vs_reference():
    ref = common.ref_backend()
    if ref is None:
        pytest.skip("oracle/_ref not present on this box")
    _ts_neutral_and_brightness_case(common.gpu_backend(), ref)


def test_run_coeval_scrolls_the_evolution_over_node_redshifts():
    """run_coeval with RECOMB_MODEL set: the chain over the node redshifts equals the explicit chain of
    single-field calls, and an output redshift between two nodes is computed from the node above it
    without feeding the evolution (coeval.py:505-512)."""
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    base = _inputs("inhomogeneous", False, "E-INTEGRAL", hii=16)
    inputs = dataclasses.replace(base, node_redshifts=REDSHIFTS)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
    pfs = [pkg.perturb_field(redshift=z, initial_conditions=ics, backend=emu) for z in REDSHIFTS]
    want = _chain(emu, inputs, ics, pfs)
    got = pkg.run_coeval(inputs=inputs, initial_conditions=ics, backend=emu)
    assert [c["redshift"] for c in got] == list(REDSHIFTS)
    for c, w in zip(got, want):
        for k in ("neutral_fraction", "z_reion", "ionisation_rate_G12", "cumulative_recombinations"):
            assert np.array_equal(getattr(c["ionized_box"], k), getattr(w, k)), (c["redshift"], k)
        assert np.isfinite(c["brightness_temp"].brightness_temp).all()
    mixed = pkg.run_coeval(out_redshifts=(8.5, 7.0), inputs=inputs, initial_conditions=ics, backend=emu)
    assert [c["redshift"] for c in mixed] == [8.5, 7.0]
    assert np.array_equal(mixed[1]["ionized_box"].cumulative_recombinations, want[2].cumulative_recombinations)
    z85 = mixed[0]["ionized_box"]
    assert set(np.unique(z85.z_reion)) <= {-1.0, 9.0, 8.5}
    # default nodes: log-spaced from the lowest output up to Z_HEAT_MAX
    nodes = pkg.get_logspaced_redshifts(7.0, 1.02, 35.0)
    assert abs(nodes[-1] - 7.0) < 1e-9 and nodes[0] >= 35.0 and all(a > b for a, b in zip(nodes, nodes[1:]))
    assert abs((1 + nodes[0]) / (1 + nodes[1]) - 1.02) < 1e-9


# ---- committed fixtures (tests/golden/recomb.npz, generator tests/golden/make_golden_recomb.py): these
# ---- run wherever the repository is, with or without oracle/_ref
def _golden():
    return np.load(common.GOLDEN / "recomb.npz")


def test_shipped_library_rate_table_matches_golden():
    """init_MHR / splined_recombination_rate are host code: the shipped CUDA library computes them
    without a GPU, and they equal what the compiled reference produced (75 000 adaptive integrals, three
    parameter splines, 300 rate splines)."""
    from test_host_logic import _host_backend
    be = _host_backend()
    g = _golden()
    inputs = _inputs("inhomogeneous", False, "E-INTEGRAL", hii=16)
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True, recomb=True)
    got = np.array([[be.lib.splined_recombination_rate(float(z), float(gm)) for gm in g["rate_gamma"]] for z in g["rate_z"]])
    assert np.allclose(got, g["rate"], rtol=1e-10, atol=0)
    be.state.free()
    assert np.isnan(be.lib.splined_recombination_rate(8.0, 1.0))  # freed tables fail loudly, not silently


@pytest.mark.parametrize("name", GOLDEN_CHAIN_CASES)
def test_emulated_chain_reproduces_golden(name):
    emu = common.emu_backend()
    if emu is None:
        pytest.skip("tests/_emu not built")
    g = _golden()
    inputs = _inputs(hii=16, **CASES[name])
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
    pfs = [pkg.perturb_field(redshift=z, initial_conditions=ics, backend=emu) for z in REDSHIFTS]
    for z, ib in zip(REDSHIFTS, _chain(emu, inputs, ics, pfs)):
        for k in GOLDEN_FIELDS:
            u, v = getattr(ib, k), g[f"{name}_{int(z)}_{k}"]
            if k == "neutral_fraction":
                assert np.array_equal(u == 0, v == 0), (name, z)
            tol = TOL_GLOBAL_NREC if v.size == 1 else 5 * common.TOL_FIELD  # ICs and density are recomputed here too
            assert np.abs(u - v).max() <= tol * max(np.abs(v).max(), 1e-30), (name, z, k)


def test_ragged_noncubic_grid_with_recombinations_and_ts_emulated():
    """The reference's awkward test sizes (35 cells, non-cubic z extent 42: mixed-radix transforms, rows that
    are not a multiple of four) through the general route -- filtered N_rec and x_e together."""
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    inp = common.make_inputs(hii=35, dim=70, seed=3)
    ao = dataclasses.replace(inp.astro_options, RECOMB_MODEL="inhomogeneous", CELL_RECOMB=False, USE_TS_FLUCT=True)
    so = dataclasses.replace(inp.simulation_options, NON_CUBIC_FACTOR=1.2)
    inp = dataclasses.replace(inp, astro_options=ao, simulation_options=so)
    ics = pkg.compute_initial_conditions(inputs=inp, backend=ref)
    pfs = [pkg.perturb_field(redshift=z, initial_conditions=ics, backend=ref) for z in (9.0, 7.5)]
    for a, b in zip(_chain(emu, inp, ics, pfs), _chain(ref, inp, ics, pfs)):
        assert a.neutral_fraction.shape == (35, 35, 42)
        assert common.compare_ionized(a, b)["mask_mismatch"] == 0
        assert np.array_equal(a.mean_free_path, b.mean_free_path)
