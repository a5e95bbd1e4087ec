#!/usr/bin/env python
"""Generate the committed golden fixtures from the compiled reference (oracle/_ref).

Run in the build container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden.py
Writes
    tests/golden/<case>.npz       inputs/outputs of ICs -> perturb -> ionize at z=8 for each case
                                  of tests/common.py::GOLDEN_CASES (float32 arrays, compressed)
    tests/golden/recfast_table.npz  the four columns of recfast_LCDM.dat (physics data the
                                  reference reads in init_heat; needed on boxes without the
                                  reference tree)
    tests/golden/host_scalars.npz  dicke / sigma_z0 / dsigmasqdm_z0 / power_in_k samples
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import common  # noqa: E402

pkg = common.pkg


def main():
    ref = common.ref_backend()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    out = Path(__file__).resolve().parent
    for name, cfg in common.GOLDEN_CASES.items():
        inputs = common.make_inputs(**cfg)
        ics = pkg.compute_initial_conditions(inputs=inputs, backend=ref)
        pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=ref)
        ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=ref)
        d = {"redshift": 8.0, "ib_mean_f_coll": ib.mean_f_coll,
             "ib_log10_Mturnover_ave": ib.log10_Mturnover_ave}
        if name == common.GOLDEN_BASE:  # ICs and the perturbed field are shared by all cases
            for k, v in ics.arrays().items():
                if k.startswith("hires_v"):
                    continue  # scratch (phi_ii) left in the hires 2LPT arrays
                d["ics_" + k] = v
            for k, v in pf.arrays().items():
                d["pf_" + k] = v
        for k, v in ib.arrays().items():
            d["ib_" + k] = v
        np.savez_compressed(out / f"{name}.npz", **d)
        print(name, "xH =", ib.global_xH)
    tab = np.loadtxt(common.table_dir() / "recfast_LCDM.dat")
    np.savez_compressed(out / "recfast_table.npz", z=tab[:, 0], xe=tab[:, 1], col3=tab[:, 2], tk=tab[:, 3])
    inputs = common.make_inputs()
    ref.state.init(inputs, broadcast_inputs=True, ps=True)
    zs = np.array([0.0, 5.0, 8.0, 12.5, 35.0, 300.0])
    Ms = np.logspace(3, 17, 15)
    ks = np.logspace(-3, 2, 11)
    np.savez(out / "host_scalars.npz", z=zs, dicke=[ref.lib.dicke(z) for z in zs], M=Ms,
             sigma=[ref.lib.sigma_z0(M) for M in Ms], dsigmasqdm=[ref.lib.dsigmasqdm_z0(M) for M in Ms],
             k=ks, power=[ref.lib.power_in_k(k) for k in ks])
    print("done")


if __name__ == "__main__":
    main()
