"""Random option combinations: the product against the compiled reference (TEST INFRASTRUCTURE, not run by pytest).

    python tools/fuzz_parity.py [--seed S] [--cases N] [--backend emu|gpu] [--spec]

Each case draws grid sizes (cubic and non-cubic, odd and even, N_THREADS 1-3), the source model (E-INTEGRAL,
CONST-ION-EFF, L-INTEGRAL with and without the exponential filter), perturbation algorithm, hi-res perturbation,
smoothing, 3-D velocities, mass function, filter, integration method, R_BUBBLE_MAX, efficiency, redshift and seed,
runs ICs -> perturb -> [halo box] -> ionize on both sides and applies the bars of tests/common.py (fields 2e-5,
velocities 2e-4, mask identical outside the threshold band).  With --spec the single-sweep ladder is also compared
bit for bit with the two-sweep ladder.  The fixed option matrix of tests/test_option_matrix.py covers each option
once; this covers their interactions.  Findings so far are listed in DESIGN.md section 2.
"""
import argparse
import os
import random
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import common  # noqa: E402

pkg = common.pkg


def draw(rng):
    hii = rng.choice([int(v) for v in os.environ["FUZZ_HII"].split(",")] if os.environ.get("FUZZ_HII")
                     else [12, 16, 20, 24, 28, 32])  # FUZZ_HII=40,48,64: larger grids (slow in the emulation)
    sim = dict(HII_DIM=hii, DIM=hii * rng.choice([2, 3]), BOX_LEN=rng.choice([1.0, 1.5, 2.0, 3.0]) * hii,
               N_THREADS=rng.choice([1, 1, 2, 3]), NON_CUBIC_FACTOR=rng.choice([1.0, 1.0, 1.0, 1.25, 1.5]))
    matter = dict(SOURCE_MODEL=rng.choice(["E-INTEGRAL", "CONST-ION-EFF", "L-INTEGRAL"]),
                  PERTURB_ALGORITHM=rng.choice(["2LPT", "2LPT", "ZELDOVICH", "LINEAR"]),
                  PERTURB_ON_HIGH_RES=rng.choice([False, False, True]),
                  SMOOTH_EVOLVED_DENSITY_FIELD=rng.choice([False, False, True]),
                  KEEP_3D_VELOCITIES=rng.choice([False, True]),
                  HMF=rng.choice(["ST", "ST", "PS", "WATSON", "WATSON-Z", "DELOS"]),
                  MINIMIZE_MEMORY=rng.choice([False, True]))
    aopt = dict(USE_EXP_FILTER=False, CELL_RECOMB=rng.choice([False, True]), USE_LYA_HEATING=False,
                USE_UPPER_STELLAR_TURNOVER=False,
                HII_FILTER=rng.choice(["spherical-tophat", "spherical-tophat", "gaussian", "sharp-k"]),
                INTEGRATION_METHOD_ATOMIC=rng.choice(["GSL-QAG", "GAUSS-LEGENDRE", "GAUSS-LEGENDRE", "GAMMA-APPROX"]))
    if matter["SOURCE_MODEL"] == "L-INTEGRAL" and aopt["HII_FILTER"] == "spherical-tophat" and rng.random() < 0.5:
        aopt.update(USE_EXP_FILTER=True, CELL_RECOMB=True)
    astro = dict(R_BUBBLE_MAX=rng.choice([10.0, 15.0, 30.0]), HII_EFF_FACTOR=rng.choice([20.0, 30.0, 50.0]),
                 DELTA_R_HII_FACTOR=rng.choice([1.1, 1.1, 1.1, 1.1, 1.05, 1.03]))  # the finer steps pass 64 radii
    cosmo = {}
    if rng.random() < 0.5:  # another cosmology / transfer function / star-formation scaling
        cosmo = dict(SIGMA_8=rng.choice([0.75, 0.8102, 0.9]), hlittle=rng.choice([0.6766, 0.7]),
                     OMm=rng.choice([0.27, 0.30966]), POWER_INDEX=rng.choice([0.95, 0.9665]))
        matter["POWER_SPECTRUM"] = rng.choice(["EH", "BBKS", "EFSTATHIOU", "PEEBLES", "WHITE"])
        astro.update(F_STAR10=rng.choice([-1.5, -1.3, -1.0]), ALPHA_STAR=rng.choice([0.3, 0.5]),
                     F_ESC10=rng.choice([-1.2, -1.0, -0.7]), ALPHA_ESC=rng.choice([-0.5, -0.2, 0.0]),
                     M_TURN=rng.choice([8.0, 8.7, 9.3]), t_STAR=rng.choice([0.3, 0.5]))
    if rng.random() < 0.12:  # tabulated transfer functions (synthetic: classy is not in the image) and v_cb fluctuations
        matter["POWER_SPECTRUM"] = "CLASS"
        matter["V_CB_MODEL"] = rng.choice(["NONE", "FLUCTS"])
        # sigma_8-normalised: with the A_s normalisation the synthetic tables give |delta| ~ 1e-4, where the float
        # rounding of (1 + delta) - 1 is all that is left to compare (tests/test_class_tables.py covers A_s on the ICs)
        cosmo["_class_sigma8"] = True
    if matter.get("POWER_SPECTRUM") in ("PEEBLES", "WHITE") and aopt["INTEGRATION_METHOD_ATOMIC"] == "GAMMA-APPROX":
        # the triple power law of sigma(M) behind the approximation does not hold for these spectra: with positive
        # exponents the sum of incomplete gamma functions cancels to noise on both sides (QAG gives ~0 there)
        aopt["INTEGRATION_METHOD_ATOMIC"] = "GAUSS-LEGENDRE"
    if rng.random() < 0.3 and not aopt["USE_EXP_FILTER"]:  # two chained snapshots with recombinations
        aopt["RECOMB_MODEL"] = rng.choice(["homogeneous", "inhomogeneous"])
        if aopt["RECOMB_MODEL"] == "homogeneous":
            aopt["CELL_RECOMB"] = True  # the reference refuses the homogeneous model with filtered recombinations
    if rng.random() < 0.2 and matter["SOURCE_MODEL"] != "L-INTEGRAL":
        aopt["USE_TS_FLUCT"] = True  # x_e filtering, T_k / T_s from a synthetic TsBox (the same on both sides)
    return dict(sim=sim, matter=matter, aopt=aopt, astro=astro, cosmo=cosmo,
                z=rng.choice([5.5, 6.5, 7.0, 8.0, 9.5, 12.0, 25.0]), seed=rng.randrange(1, 10**6))


def ladder(be, spec, **kw):
    old = os.environ.pop("B200_SPEC", None)
    if not spec:
        os.environ["B200_SPEC"] = "0"
    try:
        return pkg.compute_ionization_field(backend=be, **kw)
    finally:
        os.environ.pop("B200_SPEC", None)
        if old is not None:
            os.environ["B200_SPEC"] = old


def synthetic_ts(inputs, pf):
    """tests/test_recombinations.py::_synthetic_ts: x_e of a few per cent following the density with excursions
    beyond [0, 1], adiabatic-like T_k, T_s between it and the CMB (the spin-temperature calculation is out of scope)."""
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


class ReferenceRefused(Exception):
    """The reference itself returned an error status for this draw: nothing to compare."""


class _Ref:
    """The reference backend's calls, with its error statuses set apart from the product's."""

    def __init__(self, be):
        self.be = be

    def __getattr__(self, name):
        fn = getattr(pkg, name)

        def call(**kw):
            try:
                return fn(backend=self.be, **kw)
            except pkg.BackendError as e:
                raise ReferenceRefused(str(e), e.code) from e
        return call


def run_case(be, ref, c, spec):
    cosmo = dict(c["cosmo"])
    class_sigma8 = cosmo.pop("_class_sigma8", None)
    class_tables = None
    if c["matter"].get("POWER_SPECTRUM") == "CLASS":
        import test_class_tables as tct
        td, tv = tct._tables()
        class_tables = tct.CosmoTables(ps_norm=cosmo.get("SIGMA_8", 0.8102) if class_sigma8 else 2.1e-9,
                                       USE_SIGMA_8=bool(class_sigma8), transfer_density=td,
                                       transfer_vcb=tv if c["matter"].get("V_CB_MODEL") == "FLUCTS" else None)
    inputs = pkg.InputParameters(
        random_seed=c["seed"], cosmo_params=pkg.CosmoParams(**cosmo), class_tables=class_tables,
        simulation_options=pkg.SimulationOptions(**c["sim"]),
        matter_options=pkg.MatterOptions(**c["matter"]), astro_params=pkg.AstroParams(**c["astro"]),
        astro_options=pkg.AstroOptions(**c["aopt"]))
    z, lagrangian = c["z"], c["matter"]["SOURCE_MODEL"] == "L-INTEGRAL"
    R = _Ref(ref)
    r_ics = R.compute_initial_conditions(inputs=inputs)
    common.compare_struct(pkg.compute_initial_conditions(inputs=inputs, backend=be), r_ics)
    r_pf = R.perturb_field(redshift=z, initial_conditions=r_ics)
    pf = pkg.perturb_field(redshift=z, initial_conditions=r_ics, backend=be)
    common.compare_struct(pf, r_pf, tols={k: common.TOL_VELOCITY for k in ("velocity_x", "velocity_y", "velocity_z")})
    r_hb = None
    if lagrangian:
        r_hb = R.compute_halobox(redshift=z, initial_conditions=r_ics)
        # QAG stops at a relative tolerance of 1e-3 (hmf.c:596): a last-bit difference in the integrand can change
        # where it stops subdividing, so that is the bar for its tables, times three where the mean fix divides two such
        # integrals (seen with the PEEBLES spectrum only: 2e-4 and 2.7e-3)
        qag = c["aopt"]["INTEGRATION_METHOD_ATOMIC"] == "GSL-QAG" or c["matter"]["HMF"] not in ("PS", "ST", "DELOS")
        hb_tol = 3e-3 if qag else common.TOL_FIELD  # (the mean fix of the other mass functions is a ratio of QAG integrals)
        common.compare_struct(pkg.compute_halobox(redshift=z, initial_conditions=r_ics, backend=be), r_hb, tol=hb_tol)
    kw = dict(perturbed_field=r_pf, initial_conditions=r_ics, halobox=r_hb)
    ts_on = inputs.astro_options.USE_TS_FLUCT
    if ts_on:
        kw["spin_temp"] = synthetic_ts(inputs, r_pf)
    if inputs.evolution_required:  # the snapshot above (made by the reference) is the previous box of both sides
        zp = z + 1.0
        p_pf = R.perturb_field(redshift=zp, initial_conditions=r_ics)
        p_hb = R.compute_halobox(redshift=zp, initial_conditions=r_ics) if lagrangian else None
        p_ib = R.compute_ionization_field(
            perturbed_field=p_pf, initial_conditions=r_ics, halobox=p_hb,
            spin_temp=synthetic_ts(inputs, p_pf) if ts_on else None,
            previous_ionized_box=pkg.IonizedBox.initial(inputs), previous_perturbed_field=pkg.PerturbedField.initial(inputs))
        kw.update(previous_ionized_box=p_ib, previous_perturbed_field=p_pf)
    r_ib = R.compute_ionization_field(**kw)
    ib = ladder(be, True, **kw)
    if spec and not lagrangian and not inputs.evolution_required:
        two = ladder(be, False, **kw)
        for k, v in two.arrays().items():
            assert np.array_equal(v, ib.arrays()[k]), f"single-sweep ladder differs from two sweeps in {k}"
    mask_t, mask_r = ib.neutral_fraction == 0, r_ib.neutral_fraction == 0
    # Lagrangian sources: the collapsed fraction is the source grid over (1 + delta) (IonisationBox.c:1054-1066); in
    # nearly empty cells (a linearly evolved density is clipped at -1 + 1e-7) that division amplifies float rounding
    # by up to 1e7, flags included -- those cells are left out (tests/test_lagrangian_sources.py scales the bar)
    well = (1.0 + r_pf.density) > 0.1 if lagrangian else np.ones(mask_r.shape, bool)
    mism = int(((mask_t != mask_r) & well).sum())
    assert mism <= common.TOL_MASK_FRACTION * mask_r.size, f"mask differs in {mism} cells"
    same = mask_t == mask_r
    if "mean_free_path" in r_ib.arrays() and inputs.evolution_required:
        # a cell in the barrier's rounding band may cross one radius apart: counted like a mask mismatch
        crossing = ib.mean_free_path == r_ib.mean_free_path
        assert (~crossing).sum() <= max(2, common.TOL_MASK_FRACTION * mask_r.size), "first-crossing radius differs"
        same &= crossing
    same &= well
    # a linearly evolved density has cells clipped at -1 + 1e-7; with Lagrangian sources their amplified values also
    # enter the filtered grids of the neighbours (Gamma12, recombinations): only a coarse bar is meaningful there
    field_bar = 1e-3 if lagrangian and c["matter"]["PERTURB_ALGORITHM"] == "LINEAR" else common.TOL_FIELD
    for k, rv in r_ib.arrays().items():
        tv = ib.arrays()[k]
        if rv.shape != same.shape:
            continue
        if lagrangian and k == "neutral_fraction":  # partial ionisations: the bar follows the conditioning 1 / (1 + delta)
            d = np.abs(tv.astype(np.float64) - rv)[same] * np.minimum(1.0, 1.0 + r_pf.density[same])
            e = float(d.max()) if d.size else 0.0
        elif k == "kinetic_temperature" and same.any():
            # T = x_HI T_gas + (1 - x_HI) T_RE (IonisationBox.c:1213-1245): one ulp of a float x_HI next to 1 is
            # 6e-8 * T_RE = 1.2e-3 K whatever the scale of T_gas (34 K at z = 25)
            d = np.maximum(np.abs(tv.astype(np.float64) - rv)[same] - 2 * 6e-8 * inputs.astro_params.T_RE, 0.0)
            e = float(d.max() / max(np.abs(rv[same]).max(), 1e-300))
        else:
            e = common.rel_err(tv[same], rv[same]) if same.any() else 0.0
        assert e <= field_bar, f"{k}: rel err {e:.3e}"
    # brightness temperature of the reference's boxes (BrightnessTemperatureBox.c:22-105)
    tb = dict(ionized_box=r_ib, perturbed_field=r_pf, spin_temp=kw.get("spin_temp"))
    common.compare_struct(pkg.brightness_temperature(backend=be, **tb), R.brightness_temperature(**tb))
    # Eulerian: the analytic mean; Lagrangian: a float grid mean (IonisationBox.c:1623-1628)
    bar = 2e-6 if lagrangian else 1e-9
    assert abs(ib.mean_f_coll - r_ib.mean_f_coll) <= bar * abs(r_ib.mean_f_coll), "mean_f_coll"
    return mism, r_ib.global_xH


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cases", type=int, default=30)
    ap.add_argument("--backend", choices=["emu", "gpu"], default="emu")
    ap.add_argument("--spec", action="store_true")
    ap.add_argument("--start", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--carry", type=int, nargs=2, default=(0, 0), help=argparse.SUPPRESS)
    args = ap.parse_args()
    ref = common.ref_backend()
    be = common.emu_backend() if args.backend == "emu" else common.gpu_backend()
    if ref is None or be is None:
        sys.exit("needs oracle/_ref and the chosen product library")
    rng = random.Random(args.seed)
    ran, bad = args.carry
    for it in range(args.cases):
        c = draw(rng)
        if it < args.start:
            continue
        t = time.time()
        try:
            mism, xh = run_case(be, ref, c, args.spec)
            ran += 1
            print(f"{it:3d} ok   {time.time() - t:5.1f}s  xH={xh:.3f} mask_mismatch={mism}  {c['matter']['SOURCE_MODEL']}", flush=True)
        except AssertionError as e:
            bad += 1
            print(f"{it:3d} PARITY FAIL: {e}\n      {c}", flush=True)
        except ValueError as e:  # the input validation both sides share
            print(f"{it:3d} invalid inputs: {e}", flush=True)
        except pkg.BackendError as e:  # the reference computed it, the product returned an error status
            bad += 1
            print(f"{it:3d} PRODUCT REFUSED what the reference computes: {e}\n      {c}", flush=True)
        except ReferenceRefused as e:
            print(f"{it:3d} reference refused: {e.args[0]}\n      {c}", flush=True)
            if e.args[1] != 3:
                # the reference leaves an exception by longjmp out of an OpenMP region (a GSL error inside a table
                # build): its state is undefined afterwards, so the remaining cases run in a fresh process
                sys.stdout.flush()
                os.execv(sys.executable, [sys.executable, __file__, "--seed", str(args.seed), "--cases", str(args.cases),
                                          "--backend", args.backend, "--start", str(it + 1), "--carry", str(ran), str(bad)]
                         + (["--spec"] if args.spec else []))
        except Exception as e:  # noqa: BLE001
            bad += 1
            print(f"{it:3d} ERROR {e!r}\n      {c}", flush=True)
            traceback.print_exc()
    print(f"{ran} cases compared, {bad} failures")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
