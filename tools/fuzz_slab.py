"""Random grids and options: the slab-decomposed box against the single-rank box, bit for bit (TEST INFRASTRUCTURE).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/fuzz_slab.py [--seed S] [--cases N] [--backend emu|gpu] [--mode slab|radius|share]

Every rank draws the same cases.  Per case: whole-box ICs -> perturb -> ionize on this rank (the single-GPU path),
then the same box on x-slabs over the ranks (slab ICs, slab perturb with halo pull, slab ionize with the slab FFTs and
the barrier-kernel reductions); every output slab must equal the matching planes of the whole box exactly.  A case
the slab entry points refuse (status 3 on every rank) is reported as refused, not as a failure.  The fixed cases of
tests/test_multiprocess.py cover two grids; this covers shapes (non-cubic, mixed radix, one plane per rank) and
option combinations.  --mode radius: the other one-box partition (deposit by x-slab + all-reduce, radii split over
the ranks); --mode share: redshift-parallel ranks on a shared upload of the initial conditions.
"""
import argparse
import random
import sys
import traceback
from pathlib import Path

import numpy as np
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import common  # noqa: E402

pkg = common.pkg


def draw(rng, world):
    hii = world * rng.choice([1, 2, 3, 4, 5, 6, 8])
    while hii < 8:
        hii *= 2
    sim = dict(HII_DIM=hii, DIM=hii * rng.choice([2, 3, 4]), BOX_LEN=rng.choice([1.0, 1.5, 2.0]) * hii, N_THREADS=1,
               NON_CUBIC_FACTOR=rng.choice([1.0, 1.0, 1.25, 1.5]))
    matter = dict(SOURCE_MODEL=rng.choice(["E-INTEGRAL", "CONST-ION-EFF"]),
                  PERTURB_ALGORITHM=rng.choice(["2LPT", "2LPT", "ZELDOVICH", "LINEAR"]),
                  SMOOTH_EVOLVED_DENSITY_FIELD=rng.choice([False, False, True]))
    aopt = dict(USE_EXP_FILTER=False, CELL_RECOMB=False, USE_LYA_HEATING=False, USE_UPPER_STELLAR_TURNOVER=False,
                HII_FILTER=rng.choice(["spherical-tophat", "spherical-tophat", "gaussian", "sharp-k"]))
    astro = dict(R_BUBBLE_MAX=rng.choice([8.0, 15.0, 30.0]), HII_EFF_FACTOR=rng.choice([20.0, 30.0, 50.0]),
                 DELTA_R_HII_FACTOR=rng.choice([1.1, 1.1, 1.1, 1.04, 1.02]))  # up to ~200 radii
    return dict(sim=sim, matter=matter, aopt=aopt, astro=astro, z=rng.choice([6.0, 7.0, 8.0, 9.5, 12.0]),
                seed=rng.randrange(1, 10**6))


def run_case_radius(be, c, rank, world, device=None):
    """the other one-box partition: deposit by x-slab + all_reduce(SUM), radii split over the ranks + all_reduce(MAX)"""
    import torch
    inputs = pkg.InputParameters(
        random_seed=c["seed"], simulation_options=pkg.SimulationOptions(**c["sim"]),
        matter_options=pkg.MatterOptions(**c["matter"]), astro_params=pkg.AstroParams(**c["astro"]),
        astro_options=pkg.AstroOptions(**c["aopt"]))
    z = c["z"]
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=be)
    whole = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be)
    names = ["hires_density", "lowres_vx", "lowres_vy", "lowres_vz"]
    if c["matter"]["PERTURB_ALGORITHM"] == "2LPT":
        names += ["lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
    dev = device or "cpu"
    ppf = pkg.perturb_slab_parallel(redshift=z, ics={k: torch.from_numpy(getattr(ics, k)).to(dev) for k in names},
                                    inputs=inputs, backend=be)
    for k in ("density", "velocity_z"):
        assert np.array_equal(ppf[k].cpu().numpy(), getattr(pf, k)), f"partitioned perturb: {k}"
    part = pkg.ionize_radius_parallel(redshift=z, density=ppf["density"], inputs=inputs, backend=be)
    for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion"):
        assert np.array_equal(part[k].cpu().numpy(), getattr(whole, k).reshape(part[k].shape)), f"radius-parallel ionize: {k}"
    assert part["mean_f_coll"] == whole.mean_f_coll, "mean_f_coll"
    return whole.global_xH


def run_case_share(be, c, rank, world, device=None):
    """redshift-parallel ranks on shared initial conditions (b200_ics_share): each rank uploads 1 / world of the IC
    arrays, the rest arrives from the peers' heaps; the perturbed field of the rank's own redshift must not change"""
    inputs = pkg.InputParameters(
        random_seed=c["seed"], simulation_options=pkg.SimulationOptions(**c["sim"]),
        matter_options=pkg.MatterOptions(**c["matter"]), astro_params=pkg.AstroParams(**c["astro"]),
        astro_options=pkg.AstroOptions(**c["aopt"]))
    z = c["z"] + 0.5 * rank
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    want = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=be)
    grp = pkg.SlabGroup(inputs=inputs, backend=be, heap_bytes=pkg.SlabGroup.ics_heap_bytes(inputs))
    try:
        grp.share_ics(True)
        for _ in range(2):
            got = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=be)
            for k, v in want.arrays().items():
                assert np.array_equal(v, got.arrays()[k]), f"shared upload: {k}"
        grp.share_ics(False)
        plain = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=be)
        assert np.array_equal(plain.density, want.density), "after sharing"
    finally:
        grp.close()
    return float(want.density.std())


def run_case(be, c, rank, world, device=None):
    inputs = pkg.InputParameters(
        random_seed=c["seed"], simulation_options=pkg.SimulationOptions(**c["sim"]),
        matter_options=pkg.MatterOptions(**c["matter"]), astro_params=pkg.AstroParams(**c["astro"]),
        astro_options=pkg.AstroOptions(**c["aopt"]))
    z = c["z"]
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    pf = pkg.perturb_field(redshift=z, initial_conditions=ics, backend=be)
    whole = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=be)
    grp = pkg.SlabGroup(inputs=inputs, backend=be, ics=True)
    try:
        sics = grp.initial_conditions(device=device)
        hn = inputs.simulation_options.dim // world
        for k, t in sics.items():
            full = getattr(ics, k)
            want = full[rank * hn:(rank + 1) * hn] if k == "hires_density" else grp.lowres_slab(full)
            assert np.array_equal(t.cpu().numpy(), want), f"slab ICs: {k}"
        slab = {k: v for k, v in sics.items() if k.startswith("lowres_")}  # LINEAR reads lowres_density
        slab["hires_density"] = grp.shift_hires(sics["hires_density"])
        ppf = grp.perturb(redshift=z, ics_slab=slab)
        for k in ("density", "velocity_z"):
            assert np.array_equal(ppf[k].cpu().numpy(), grp.lowres_slab(getattr(pf, k))), f"slab perturb: {k}"
        part = grp.ionize(redshift=z, density_slab=ppf["density"])
        for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion"):
            want = grp.lowres_slab(getattr(whole, k).reshape(pf.density.shape))
            assert np.array_equal(part[k].cpu().numpy(), want), f"slab ionize: {k}"
        assert part["mean_f_coll"] == whole.mean_f_coll, "mean_f_coll"
    finally:
        grp.close()
    return whole.global_xH


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cases", type=int, default=20)
    ap.add_argument("--backend", choices=["emu", "gpu"], default="emu")
    ap.add_argument("--mode", choices=["slab", "radius", "share"], default="slab")
    args = ap.parse_args()
    device = None
    if args.backend == "gpu":  # one process per GPU
        import os

        import torch
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = f"cuda:{local}"
    dist.init_process_group("gloo" if args.backend == "emu" else "nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    be = common.emu_backend() if args.backend == "emu" else common.gpu_backend()
    if device:
        assert be.lib.b200_set_device(int(device.split(":")[1])) == 0
    rng = random.Random(args.seed)
    bad = ran = 0
    for it in range(args.cases):
        c = draw(rng, world)
        try:
            xh = {"slab": run_case, "radius": run_case_radius, "share": run_case_share}[args.mode](be, c, rank, world, device)
            ran += 1
            if rank == 0:
                print(f"{it:3d} ok   xH={xh:.3f}  {c['sim']['HII_DIM']}/{c['sim']['DIM']} x{c['sim']['NON_CUBIC_FACTOR']}", flush=True)
        except AssertionError as e:
            bad += 1
            print(f"{it:3d} rank {rank} SLAB != WHOLE: {e}\n      {c}", flush=True)
        except (ValueError, pkg.BackendError) as e:
            if rank == 0:
                print(f"{it:3d} refused: {e}\n      {c}", flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print(f"{it:3d} rank {rank} ERROR {e!r}\n      {c}", flush=True)
            traceback.print_exc()
    if rank == 0:
        print(f"{ran} cases compared on {world} ranks, {bad} failures on rank 0")
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
