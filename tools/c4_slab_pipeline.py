#!/usr/bin/env python
"""BASELINE config C4 on the GPUs of one node: a coeval box whose hi-res grid does not fit one GPU, end to end on
x-slabs -- initial conditions, perturbed field and ionized box, no whole box anywhere.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
        tools/c4_slab_pipeline.py --hii-dim 1024 --dim 2048 --box-len 1000

Prints one JSON line on rank 0: device milliseconds of the three library calls (max over ranks), the global neutral
fraction and a few sanity numbers.  ICs use the counter-based field (B200_IC_RNG=device): walking the reference's
sequential Gaussian stream at DIM >= 2048 takes minutes and is not what is measured here."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import common  # noqa: E402

pkg = importlib.import_module("21cmfast_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--hii-dim", type=int, default=1024)
ap.add_argument("--dim", type=int, default=2048)
ap.add_argument("--box-len", type=float, default=1000.0)
ap.add_argument("--redshift", type=float, default=8.0)
ap.add_argument("--r-bubble-max", type=float, default=15.0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--emu", action="store_true", help="dry run of the script's logic on the CPU emulation (gloo)")
args = ap.parse_args()

local = int(os.environ.get("LOCAL_RANK", "0"))
if args.emu:
    dist.init_process_group("gloo")
    dev = torch.device("cpu")
    be = common.emu_backend()
else:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    be = pkg.get_backend()
    be.set_table_path(common.table_dir())
    assert be.lib.b200_set_device(local) == 0
rank, world = dist.get_rank(), dist.get_world_size()
os.environ["B200_IC_RNG"] = "device"
ncpu = max(1, (os.cpu_count() or 1) // world)
inputs = common.make_inputs(hii=args.hii_dim, dim=args.dim, box_len=args.box_len, source="E-INTEGRAL", seed=4321,
                            n_threads=ncpu, R_BUBBLE_MAX=args.r_bubble_max)
so = inputs.simulation_options


def sync():
    if dev.type == "cuda":
        torch.cuda.synchronize()


def stats():
    a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_double()
    be.lib.b200_last_call_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    return d.value


def mx(v):
    t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def total(t):
    s = t.double().sum()
    dist.all_reduce(s)
    return float(s.item())


grp = pkg.SlabGroup(inputs=inputs, backend=be, ics=True)
rec = []
for rep in range(args.reps):
    sync()
    t0 = time.perf_counter()
    ics = grp.initial_conditions(device=dev)
    ms_ic = stats()
    sl = {k: v for k, v in ics.items() if k.startswith("lowres_v")}
    sl["hires_density"] = grp.shift_hires(ics["hires_density"])
    sync()
    t1 = time.perf_counter()
    pf = grp.perturb(redshift=args.redshift, ics_slab=sl)
    ms_pf = stats()
    ib = grp.ionize(redshift=args.redshift, density_slab=pf["density"], want_nion=False)
    ms_ib = stats()
    sync()
    t2 = time.perf_counter()
    rec.append(dict(ms_ics=mx(ms_ic), ms_perturb=mx(ms_pf), ms_ionize=mx(ms_ib), wall_ics_s=mx(t1 - t0),
                    wall_perturb_ionize_s=mx(t2 - t1)))
N = args.hii_dim ** 3
out = {"config": f"C4: HII_DIM={args.hii_dim} DIM={args.dim} BOX_LEN={args.box_len:g} z={args.redshift} E-INTEGRAL, "
                 f"{world} GPUs, everything on x-slabs (ICs, perturb, ionize)",
       "global_xH": total(ib["neutral_fraction"]) / N,
       "mean_density": total(pf["density"]) / N, "rms_hires_density": (total(ics["hires_density"] ** 2) / args.dim ** 3) ** 0.5,
       "mean_f_coll": ib["mean_f_coll"], "runs": rec,
       "cells_per_s_perturb_ionize": N / (1e-3 * (rec[-1]["ms_perturb"] + rec[-1]["ms_ionize"])),
       "torch_hbm_per_rank_GB": (torch.cuda.max_memory_allocated(dev) / 1e9) if dev.type == "cuda" else None}
grp.close()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
