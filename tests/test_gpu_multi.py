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


WORKER_SLAB = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
ndev = torch.cuda.device_count()
local = int(os.environ["LOCAL_RANK"]) % ndev
torch.cuda.set_device(local)
world = int(os.environ["WORLD_SIZE"])
if world <= ndev:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo")      # several ranks share one GPU: the plumbing runs on the host
be = pkg.get_backend(); be.set_table_path(common.table_dir())
assert be.lib.b200_set_device(local) == 0
dev = torch.device("cuda", local)
for hii, dim, kw in {cases}:
    inputs = common.make_inputs(hii=hii, dim=dim, seed=777, **kw)
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=be)
    whole = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be)
    grp = pkg.SlabGroup(inputs=inputs, backend=be, ics=True)
    sics = grp.initial_conditions(device=dev)             # slab-decomposed ICs: bit-identical slabs
    hn = inputs.simulation_options.dim // world
    rk = dist.get_rank()
    for k, t in sics.items():
        full = getattr(ics, k)
        want = full[rk * hn:(rk + 1) * hn] if k == "hires_density" else grp.lowres_slab(full)
        assert np.array_equal(t.cpu().numpy(), want), (hii, k, float(np.abs(t.cpu().numpy() - want).max()))
    lo = ["lowres_vx", "lowres_vy", "lowres_vz", "lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
    slab = {{k: torch.from_numpy(np.ascontiguousarray(grp.lowres_slab(getattr(ics, k)))).to(dev) for k in lo
            if getattr(ics, k) is not None}}
    slab["hires_density"] = torch.from_numpy(grp.hires_slab(ics.hires_density)).to(dev)
    for rep in range(2):                                  # twice: heap reuse across calls
        ppf = grp.perturb(redshift=8.0, ics_slab=slab)
        for k in ("density", "velocity_z"):
            a, b = ppf[k].cpu().numpy(), grp.lowres_slab(getattr(pf, k))
            assert np.array_equal(a, b), (hii, k, float(np.abs(a - b).max()))
        part = grp.ionize(redshift=8.0, density_slab=ppf["density"])
        for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion"):
            a, b = part[k].cpu().numpy(), grp.lowres_slab(getattr(whole, k).reshape(pf.density.shape))
            assert np.array_equal(a, b), (hii, k, float(np.abs(a - b).max()))
        assert part["mean_f_coll"] == whole.mean_f_coll
    grp.close()
    if dist.get_rank() == 0: print("OK slab", world, hii, float(whole.neutral_fraction.mean()))
dist.destroy_process_group()
'''

SLAB_CASES = [(64, 128, {}), (128, 384, {}), (48, 96, dict(perturb="ZELDOVICH", hii_filter="sharp-k", source="CONST-ION-EFF"))]


def _run_slab(tmp_path, nproc, port):
    script = tmp_path / "worker_slab.py"
    script.write_text(WORKER_SLAB.format(root=ROOT, cases=repr(SLAB_CASES)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ))
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("OK slab") == len(SLAB_CASES)


def test_slab_decomposed_box_one_rank(tmp_path):
    """The slab code path on ONE GPU (the rank is its own peer): scatter stores, halo pull, barrier kernel
    and the gathered reductions, bit-identical to the plain entry points."""
    _run_slab(tmp_path, 1, 29561)


def test_slab_decomposed_box_two_ranks(tmp_path):
    """Two ranks -- on two GPUs over NVLink where the box has them, else sharing the one GPU (the peer heap
    is then a CUDA-IPC mapping of memory on the same device): every output slab bit-identical to the
    single-GPU box."""
    _run_slab(tmp_path, 2, 29563)


def test_slab_decomposed_box_all_gpus(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip("needs at least four GPUs")
    _run_slab(tmp_path, 8 if n >= 8 else 4, 29565)
