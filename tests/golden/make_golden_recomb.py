#!/usr/bin/env python
"""Golden fixtures of the recombination / evolution path from the compiled reference (oracle/_ref).

Run in the build container:
    python tests/golden/make_golden_recomb.py
Writes tests/golden/recomb.npz:
    rate_z, rate_gamma, rate     splined_recombination_rate(z, Gamma12) after init_MHR
    <case>_<z>_<field>           IonizedBox fields of the chained snapshots z = 9 -> 8 -> 7 at 16^3 for the
                                 cases of tests/test_recombinations.py::GOLDEN_CHAIN_CASES
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import common  # noqa: E402
import test_recombinations as tr  # noqa: E402

pkg = common.pkg


def main():
    ref = common.ref_backend()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    d = {}
    inputs = tr._inputs("inhomogeneous", False, "E-INTEGRAL", hii=16)
    ref.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True, recomb=True)
    d["rate_z"], d["rate_gamma"] = tr.RATE_Z, tr.RATE_GAMMA
    d["rate"] = np.array([[ref.lib.splined_recombination_rate(float(z), float(g)) for g in tr.RATE_GAMMA] for z in tr.RATE_Z])
    for name in tr.GOLDEN_CHAIN_CASES:
        kw = tr.CASES[name]
        inputs = tr._inputs(hii=16, **kw)
        ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
        pfs = [pkg.perturb_field(redshift=z, initial_conditions=ics, backend=ref) for z in tr.REDSHIFTS]
        for z, ib in zip(tr.REDSHIFTS, tr._chain(ref, inputs, ics, pfs)):
            for k in tr.GOLDEN_FIELDS:
                d[f"{name}_{int(z)}_{k}"] = getattr(ib, k)
            print(name, z, "xH =", ib.global_xH)
    np.savez_compressed(Path(__file__).resolve().parent / "recomb.npz", **d)


if __name__ == "__main__":
    main()
