"""Shared helpers for the test-suite, bench.py and __graft_entry__.smoke().

Only test infrastructure lives here: building inputs, locating the physics data table, binding the
checkers (compiled reference in oracle/_ref, host-emulation library in tests/_emu) and comparing
output structs with the tolerances stated in DESIGN.md.
"""
from __future__ import annotations

import importlib
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
pkg = importlib.import_module("21cmfast_b200")

# float32 tolerances (relative to the field's max |value|), see DESIGN.md "Parity definition"
TOL_FIELD = 2e-5      # every float32 field except velocities
TOL_VELOCITY = 2e-4   # velocities inherit the reference's 1e-10 finite-difference noise in dD/dt
TOL_MASK_FRACTION = 1e-4  # fraction of cells whose ionised flag may differ (threshold band)


def make_inputs(hii=32, dim=None, box_len=None, source="E-INTEGRAL", hii_filter="spherical-tophat",
                seed=12345, perturb="2LPT", n_threads=1, smooth_evolved=False, **astro):
    dim = dim or 2 * hii
    box_len = box_len or 1.5 * hii
    return pkg.InputParameters(
        random_seed=seed,
        simulation_options=pkg.SimulationOptions(HII_DIM=hii, DIM=dim, BOX_LEN=box_len, N_THREADS=n_threads),
        matter_options=pkg.MatterOptions(SOURCE_MODEL=source, PERTURB_ALGORITHM=perturb,
                                         SMOOTH_EVOLVED_DENSITY_FIELD=smooth_evolved),
        astro_params=pkg.AstroParams(**astro),
        astro_options=pkg.AstroOptions(USE_EXP_FILTER=False, CELL_RECOMB=False, USE_LYA_HEATING=False,
                                       USE_UPPER_STELLAR_TURNOVER=False, HII_FILTER=hii_filter),
    )


def table_dir() -> Path:
    """Directory holding recfast_LCDM.dat (what config_settings.external_table_path points at): the
    reference's own data directory where it is present, else the table packaged with the product."""
    for p in (os.environ.get("PY21CMFAST_DATA"), ROOT / "oracle" / "_ref" / "data",
              "/root/reference/src/py21cmfast/_data"):
        if p and Path(p, "recfast_LCDM.dat").exists():
            return Path(p)
    return importlib.import_module("21cmfast_b200._data").default_table_dir()


def ref_backend():
    """The compiled reference (oracle/_ref) or None when it did not travel with the repo."""
    from oracle import ref_harness
    if not ref_harness.available():
        return None
    return ref_harness.ref_backend()


_emu = None


def emu_backend():
    global _emu
    path = ROOT / "tests" / "_emu" / "libb200_emu.so"
    if not path.exists():
        return None
    if _emu is None:
        _emu = pkg.Backend(path)
        _emu.set_table_path(table_dir())
    return _emu


def gpu_backend():
    be = pkg.get_backend()  # raises ImportError loudly if the CUDA library is missing
    be.set_table_path(table_dir())
    return be


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def compare_struct(test, ref, tol=TOL_FIELD, tols=None, skip=()):
    """Return {field: rel_err}; raise AssertionError on the first field over tolerance."""
    out = {}
    for k, rv in ref.arrays().items():
        if k in skip:
            continue
        tv = test.arrays().get(k)
        assert tv is not None, f"missing array {k}"
        assert tv.shape == rv.shape, (k, tv.shape, rv.shape)
        e = rel_err(tv, rv)
        out[k] = e
        lim = (tols or {}).get(k, tol)
        assert e <= lim, f"{k}: rel err {e:.3e} > {lim:.1e}"
    return out


def compare_ionized(test, ref):
    """IonizedBox comparison: float fields within TOL_FIELD, ionised mask bit-exact outside the
    threshold band (reported as a mismatch fraction, bounded by TOL_MASK_FRACTION)."""
    mask_t, mask_r = test.neutral_fraction == 0, ref.neutral_fraction == 0
    mism = int((mask_t != mask_r).sum())
    frac = mism / mask_r.size
    assert frac <= TOL_MASK_FRACTION, f"ionised mask differs in {mism} cells ({frac:.2e})"
    same = mask_t == mask_r
    out = {"mask_mismatch": mism}
    for k, rv in ref.arrays().items():
        tv = test.arrays()[k]
        if rv.size == 1:  # the homogeneous model's single cumulative_recombinations value: checked by its own test
            continue
        if rv.shape != same.shape:
            rv, tv = rv[0], tv[0]
        e = rel_err(tv[same], rv[same]) if same.any() else 0.0
        out[k] = e
        assert e <= TOL_FIELD, f"{k}: rel err {e:.3e}"
    assert abs(test.mean_f_coll - ref.mean_f_coll) <= 1e-9 * abs(ref.mean_f_coll)
    assert abs(test.log10_Mturnover_ave - ref.log10_Mturnover_ave) <= 1e-12
    return out


GOLDEN_BASE = "einteg32"  # the case whose file also stores the (shared) ICs and perturbed field
GOLDEN_CASES = {
    "einteg32": dict(hii=32, dim=64, source="E-INTEGRAL"),
    "const32": dict(hii=32, dim=64, source="CONST-ION-EFF"),
}


def load_golden(name):
    """(inputs, ics, pf, ib) rebuilt from tests/golden/<name>.npz (generated by make_golden.py)."""
    g = dict(np.load(GOLDEN / f"{GOLDEN_BASE}.npz"))
    g.update(dict(np.load(GOLDEN / f"{name}.npz")))
    inputs = make_inputs(**GOLDEN_CASES[name])
    ics = pkg.InitialConditions.new(inputs)
    pf = pkg.PerturbedField.new(inputs, redshift=float(g["redshift"]))
    ib = pkg.IonizedBox.new(inputs, redshift=float(g["redshift"]))
    for obj, prefix in ((ics, "ics_"), (pf, "pf_"), (ib, "ib_")):
        for k in obj._arrays:
            key = prefix + k
            if key in g:
                getattr(obj, k)[...] = g[key].reshape(getattr(obj, k).shape)
            elif prefix == "ics_" and getattr(obj, k) is not None and k.startswith("hires_v"):
                setattr(obj, k, None)  # scratch outputs are not stored in the golden file
    ib.mean_f_coll = float(g["ib_mean_f_coll"])
    ib.log10_Mturnover_ave = float(g["ib_log10_Mturnover_ave"])
    return inputs, ics, pf, ib


def check_against_oracle(inputs, ics, pf, ib):
    """Compare a product run with the strongest checker available; returns its name."""
    ref = ref_backend()
    if ref is not None:
        r_ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
        compare_struct(ics, r_ics)
        r_pf = pkg.perturb_field(redshift=pf.redshift, initial_conditions=r_ics, backend=ref)
        compare_struct(pf, r_pf, tols={"velocity_z": TOL_VELOCITY})
        r_ib = pkg.compute_ionization_field(perturbed_field=r_pf, initial_conditions=r_ics, backend=ref)
        stats = compare_ionized(ib, r_ib)
        return f"oracle/_ref (mask mismatches: {stats['mask_mismatch']})"
    return "none (oracle/_ref not present; golden fixtures are exercised by tests/)"
