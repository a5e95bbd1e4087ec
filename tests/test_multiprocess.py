"""N > 1 host logic on CPU: the path shards by independent coeval boxes (one per rank, no data
path collective); the only exchange is the max-over-ranks timing reduction bench.py does.  Two
gloo ranks each run the host-emulation pipeline on their own seed and agree on the reduction."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emu = common.emu_backend()
inputs = common.make_inputs(hii=16, dim=32, seed=100 + rank)
ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=emu)
xh = torch.tensor([ib.global_xH], dtype=torch.float64)
allx = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
dist.all_gather(allx, xh)
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
vals = [x.item() for x in allx]
assert all(0 < v < 1 for v in vals) and len(set(round(v, 9) for v in vals)) == world, vals
cells = torch.tensor([16.0**3]); dist.all_reduce(cells)
assert cells.item() == world * 16**3
if rank == 0: print("OK", vals)
dist.destroy_process_group()
'''


def test_two_rank_gloo_independent_boxes(tmp_path):
    if not (ROOT / "tests" / "_emu" / "libb200_emu.so").exists():
        pytest.skip("tests/_emu not built")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "OK" in r.stdout


WORKER_RADIUS = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emu = common.emu_backend()
inputs = common.make_inputs(hii=32, dim=64, seed=4242)            # the SAME box on every rank
ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
names = ["hires_density", "lowres_vx", "lowres_vy", "lowres_vz", "lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
ppf = pkg.perturb_slab_parallel(redshift=8.0, ics={{k: torch.from_numpy(getattr(ics, k)) for k in names}},
                                inputs=inputs, backend=emu)
for k in ("density", "velocity_z"):                                 # slab deposit + all_reduce(SUM): bit-identical
    assert np.array_equal(ppf[k].numpy(), getattr(pf, k)), k
whole = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=emu)
part = pkg.ionize_radius_parallel(redshift=8.0, density=ppf["density"], inputs=inputs, backend=emu)
for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion"):
    a, b = part[k].numpy(), getattr(whole, k).reshape(part[k].shape)
    assert np.array_equal(a, b), (k, np.abs(a - b).max())          # bit-identical to the one-rank ladder
assert part["mean_f_coll"] == whole.mean_f_coll
xh = float(part["neutral_fraction"].mean())
assert 0.0 < xh < 1.0
if rank == 0: print("OK radius-parallel", world, xh)
dist.destroy_process_group()
'''


def test_two_rank_gloo_radius_parallel_box(tmp_path):
    """One box, radii split over two ranks, masks merged with all_reduce(MAX): bit-identical to the
    single-rank ladder (the partition and the merge are the N > 1 host logic of SURVEY section 8e)."""
    if not (ROOT / "tests" / "_emu" / "libb200_emu.so").exists():
        pytest.skip("tests/_emu not built")
    script = tmp_path / "worker_radius.py"
    script.write_text(WORKER_RADIUS.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "OK radius-parallel" in r.stdout


WORKER_SLAB = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emu = common.emu_backend()
for hii, dim, kw in {cases}:
    inputs = common.make_inputs(hii=hii, dim=dim, seed=4242, **kw)   # the SAME box on every rank
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
    pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
    whole = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=emu)
    grp = pkg.SlabGroup(inputs=inputs, backend=emu, ics=True)
    # slab-decomposed initial conditions (SURVEY 8e row 4): every rank walks the one Gaussian stream and keeps its
    # y-slab of the modes, the Hermitian planes are fixed on gathered copies -- bit-identical slabs
    sics = grp.initial_conditions()
    hn = inputs.simulation_options.dim // world
    for k, t in sics.items():
        full = getattr(ics, k)
        want = full[rank * hn:(rank + 1) * hn] if k == "hires_density" else grp.lowres_slab(full)
        assert np.array_equal(t.numpy(), want), (hii, k, float(np.abs(t.numpy() - want).max()))
    lo = ["lowres_density", "lowres_vx", "lowres_vy", "lowres_vz", "lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
    # the whole pipeline on slabs: the ICs made slab by slab feed the slab perturb (half-cell shift of the hi-res planes
    # through one small all-gather), which feeds the slab ionize -- no whole box anywhere
    slab = {{k: sics[k] for k in lo if k in sics}}
    slab["hires_density"] = grp.shift_hires(sics["hires_density"])
    assert np.array_equal(slab["hires_density"].numpy(), grp.hires_slab(ics.hires_density))
    ppf = grp.perturb(redshift=8.0, ics_slab=slab)
    for k in ("density", "velocity_z"):                  # slab deposit + halo pull + slab FFTs: bit-identical
        a, b = ppf[k].numpy(), grp.lowres_slab(getattr(pf, k))
        assert np.array_equal(a, b), (hii, k, float(np.abs(a - b).max()))
    part = grp.ionize(redshift=8.0, density_slab=ppf["density"])
    for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion"):
        a, b = part[k].numpy(), grp.lowres_slab(getattr(whole, k).reshape(pf.density.shape))
        assert np.array_equal(a, b), (hii, k, float(np.abs(a - b).max()))
    assert part["mean_f_coll"] == whole.mean_f_coll
    xh = float(part["neutral_fraction"].mean())
    grp.close()
    if rank == 0: print("OK slab", world, hii, xh)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("nproc", [1, 2, 4])
def test_gloo_slab_decomposed_box(tmp_path, nproc):
    """One box on x-slabs over 1 / 2 / 4 ranks: slab deposit with halo pull, slab-decomposed FFTs whose
    transposes are stores into the peers' (shared-memory) heaps, per-radius extrema / plane sums through
    the barrier kernel -- every output slab bit-identical to the single-rank box.  Power-of-two and
    mixed-radix grids, 2LPT, Zel'dovich and the linear field, top-hat (window rows), sharp-k and Gaussian."""
    if not (ROOT / "tests" / "_emu" / "libb200_emu.so").exists():
        pytest.skip("tests/_emu not built")
    cases = [(32, 64, {}), (24, 72, dict(perturb="ZELDOVICH", hii_filter="sharp-k", source="CONST-ION-EFF")),
             (16, 32, dict(perturb="LINEAR", hii_filter="gaussian"))]  # LINEAR: the rank's planes of lowres_density
    script = tmp_path / "worker_slab.py"
    script.write_text(WORKER_SLAB.format(root=ROOT, cases=repr(cases)))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29551 + nproc), str(script)],
                       capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("OK slab") == len(cases)


WORKER_ASYM = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emu = common.emu_backend()
inputs = common.make_inputs(hii=32, dim=64, seed=4242)
ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
pf = pkg.perturb_field(redshift=8.0, initial_conditions=ics, backend=emu)
whole = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=emu)
grp = pkg.SlabGroup(inputs=inputs, backend=emu)
dens = torch.from_numpy(np.ascontiguousarray(grp.lowres_slab(pf.density)))
if rank == 1:
    os.environ["B200_SPEC_QCAP"] = "3"      # only this rank's queue segments overflow
part = grp.ionize(redshift=8.0, density_slab=dens)
for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion"):
    assert np.array_equal(part[k].numpy(), grp.lowres_slab(getattr(whole, k).reshape(pf.density.shape))), k
grp.close()
print("OK asym", rank)
dist.destroy_process_group()
'''


def test_gloo_slab_ranks_leave_the_speculative_ladder_together(tmp_path):
    """Single-sweep ladder on slabs: a queue segment that overflows on ONE rank only (forced here through that
    rank's B200_SPEC_QCAP) must send every rank into the two-sweep re-run -- the failed flag is gathered over the
    ranks before the decision (a local decision left the ranks in different barrier sequences: deadlock).
    Outputs stay bit-identical to the single-rank box."""
    if not (ROOT / "tests" / "_emu" / "libb200_emu.so").exists():
        pytest.skip("tests/_emu not built")
    script = tmp_path / "worker_asym.py"
    script.write_text(WORKER_ASYM.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    env.pop("B200_SPEC_QCAP", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("OK asym") == 2


WORKER_SHARE = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emu = common.emu_backend()
inputs = common.make_inputs(hii=16, dim=48, seed=99)          # the SAME initial conditions on every rank
ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
z = 7.0 + rank                                                 # one redshift per rank
want = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=emu)
grp = pkg.SlabGroup(inputs=inputs, backend=emu, heap_bytes=pkg.SlabGroup.ics_heap_bytes(inputs))
grp.share_ics(True)
import ctypes as C
for rep in range(2):                                           # twice: the heap is carved anew by every call
    got = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=emu)
    a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_double()
    emu.lib.b200_last_call_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    for k, v in want.arrays().items():
        assert np.array_equal(v, got.arrays()[k]), (rank, k)
    n_ic = 4 * (48**3 + 6 * 16**3)
    assert abs(b.value - n_ic / world) <= 64, (b.value, n_ic / world)   # this rank uploaded its share only
grp.share_ics(False)
plain = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=emu)
assert np.array_equal(plain.density, want.density)
grp.close()
if rank == 0: print("OK shared ics", world)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("nproc", [2, 4])
def test_gloo_shared_initial_conditions(tmp_path, nproc):
    """Redshift-parallel ranks on shared ICs: every rank uploads 1 / world of the IC arrays and
    receives the rest from its peers' heaps; the perturbed field of each rank's redshift is bit-identical to the
    plain call, and the rank's host-to-device bytes are its share only."""
    if not (ROOT / "tests" / "_emu" / "libb200_emu.so").exists():
        pytest.skip("tests/_emu not built")
    script = tmp_path / "worker_share.py"
    script.write_text(WORKER_SHARE.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29571 + nproc), str(script)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "OK shared ics" in r.stdout


WORKER_COEVAL = r'''
import os, sys
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
import numpy as np, torch, torch.distributed as dist
import common
pkg = common.pkg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
emu = common.emu_backend()
inputs = common.make_inputs(hii=16, dim=32, seed=7)
ics = pkg.compute_initial_conditions(inputs=inputs, backend=emu)
zs = (9.0, 8.0, 7.0)                                    # three redshifts on two ranks: the last round has an idle rank
mine = pkg.run_coeval_parallel(out_redshifts=zs, inputs=inputs, initial_conditions=ics, backend=emu)
serial = {{r["redshift"]: r for r in pkg.run_coeval(out_redshifts=zs, inputs=inputs, initial_conditions=ics, backend=emu)}}
want = sorted(zs, reverse=True)[rank::world]
assert [r["redshift"] for r in mine] == want, (rank, [r["redshift"] for r in mine], want)
for r in mine:
    s = serial[r["redshift"]]
    for key in ("perturbed_field", "ionized_box", "brightness_temp"):
        for k, v in s[key].arrays().items():
            assert np.array_equal(v, r[key].arrays()[k]), (rank, r["redshift"], key, k)
n = torch.tensor([float(len(mine))]); dist.all_reduce(n)
assert n.item() == len(zs)
if rank == 0: print("OK coeval parallel", world, want)
dist.destroy_process_group()
'''


def test_gloo_redshift_parallel_coeval(tmp_path):
    """run_coeval_parallel: the redshifts of a run round-robin over two ranks on shared initial conditions, every box
    identical to the serial run_coeval's."""
    if not (ROOT / "tests" / "_emu" / "libb200_emu.so").exists():
        pytest.skip("tests/_emu not built")
    script = tmp_path / "worker_coeval.py"
    script.write_text(WORKER_COEVAL.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29581", str(script)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "OK coeval parallel" in r.stdout
