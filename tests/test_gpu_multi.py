"""Two-GPU test of the radius-parallel ionisation (NCCL all-reduce of the mask): the merged result
equals the single-GPU ladder bit for bit.  Skipped on boxes with one GPU."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
be = pkg.get_backend(); be.set_table_path(common.table_dir())
assert be.lib.b200_set_device(local) == 0
inputs = common.make_inputs(hii=64, dim=128, seed=777)
ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=be)
whole = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be)
names = ["hires_density", "lowres_vx", "lowres_vy", "lowres_vz", "lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
ppf = pkg.perturb_slab_parallel(redshift=8.0, ics={{k: torch.from_numpy(getattr(ics, k)).cuda() for k in names}},
                                inputs=inputs, backend=be)
for k in ("density", "velocity_z"):
    assert np.array_equal(ppf[k].cpu().numpy(), getattr(pf, k)), k
part = pkg.ionize_radius_parallel(redshift=8.0, density=ppf["density"], inputs=inputs, backend=be)
for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion"):
    a, b = part[k].cpu().numpy(), getattr(whole, k).reshape(part[k].shape)
    assert np.array_equal(a, b), (k, float(np.abs(a - b).max()))
if dist.get_rank() == 0: print("OK nccl radius-parallel", float(part["neutral_fraction"].mean()))
dist.destroy_process_group()
'''


def test_two_gpu_radius_parallel(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ))
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "OK nccl radius-parallel" in r.stdout
