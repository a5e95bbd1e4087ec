"""Extract the reference's OWN golden vectors for the scoped path into tests/golden/upstream_goldens.npz.

Run in the build container (the GPU box has no /root/reference):  python tests/golden/make_upstream_goldens.py

Source: /root/reference/tests/test_data/*.h5, the files the reference's tests/test_integration_features.py compares
against (made upstream by tests/produce_integration_test_data.py, on the upstream maintainers' machines with the real
GSL / FFTW -- nothing of this repo took part).  Kept: the four perturb-field files (PDF and P(k) of density and
velocity_z at z = 10) and the coeval spectra (z = 18) of the option sets that stay inside the scoped path --
no spin temperature, no halo catalogue, no photon conservation.  Of the lightcone groups only the
global signals (means of the coeval boxes at the node redshifts) are kept: the lightcone interpolation is post-processing
outside the path.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import h5mini  # noqa: E402

SRC = Path("/root/reference/tests/test_data")
PT = ["simple", "no2lpt", "linear", "highres"]
COEVAL = ["simple", "no-mdz", "fftw_wisdom", "fixed_halogrids", "homo", "inhomo"]

out = {}
for name in PT:
    for key, arr in h5mini.read_datasets((SRC / f"perturb_field_data_{name}.h5").read_bytes()).items():
        out[f"pt/{name}/{key}"] = arr
for name in COEVAL:
    for key, arr in h5mini.read_datasets((SRC / f"power_spectra_{name}.h5").read_bytes()).items():
        if key.startswith("coeval/"):
            out[f"coeval/{name}/{key.split('/', 1)[1]}"] = arr
        elif key.startswith("lightcone/global_"):  # means of the coeval boxes at the node redshifts
            out[f"lightcone/{name}/{key.split('/', 1)[1]}"] = arr
np.savez_compressed(HERE / "upstream_goldens.npz", **out)
print(len(out), "arrays,", (HERE / "upstream_goldens.npz").stat().st_size, "bytes")
