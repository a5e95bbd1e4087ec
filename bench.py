#!/usr/bin/env python
"""bench.py -- coeval cells/sec for perturb + ionize at one redshift (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own C, host cores

One "step" = ComputePerturbedField + ComputeIonizedBox for one coeval box with the initial
conditions already generated (ICs are outside the metric).

Default workload = the configuration BASELINE.json's metric is quoted on: z = 8, HII_DIM = 512,
DIM = 1536 (the reference's default 3x), BOX_LEN = 768 Mpc, R_BUBBLE_MAX = 40 -> 40 filter radii.

b200 arm, per rank (one process per GPU; --partition boxes: independent boxes -> weak scaling, no
collective on the data path; --partition radius: ONE box split over the GPUs, two collectives):
  value : steps with every input/output already resident in HBM (device-pointer entry points),
          timed with CUDA events on the library's stream, max over ranks.
  e2e   : the same steps through the reference-facing C-ABI (ComputePerturbedField /
          ComputeIonizedBox) with pinned HOST buffers; H2D of the initial conditions and D2H of
          every output box are inside the timed region (IC device cache disabled).
  roofline : dominant kernel by summed CUDA-event time inside the timed steps; achieved =
          algorithmic bytes per launch / mean launch time (DESIGN.md "Algorithmic bytes").
  cpu_baseline : oracle/_ref (the reference's C sources compiled here against the FFTW/GSL
          shims) on a bounded sample of the same workload, on this host's cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hii-dim", type=int, default=512)
    ap.add_argument("--dim", type=int, default=0, help="hi-res grid (default 3 x HII_DIM)")
    ap.add_argument("--box-len", type=float, default=0.0, help="Mpc (default 1.5 Mpc per cell: 768 at HII_DIM=512)")
    ap.add_argument("--redshift", type=float, default=8.0)
    ap.add_argument("--source", default="E-INTEGRAL", choices=["E-INTEGRAL", "CONST-ION-EFF"])
    ap.add_argument("--r-bubble-max", type=float, default=40.0)
    ap.add_argument("--ref-hii-dim", type=int, default=256, help="bounded CPU sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-yardstick", action="store_true", help="skip the cuFFT library-baseline leg")
    ap.add_argument("--partition", default="boxes", choices=["boxes", "radius", "slab"],
                    help="N > 1: 'boxes' = one independent coeval box per GPU (weak scaling, no collective; the line "
                         "also carries a `strong` record with both one-box partitions); 'slab' = ONE box on x-slabs: "
                         "slab-decomposed FFTs whose transposes are peer stores over NVLink, slab-local deposit with a "
                         "halo pull (strong scaling); 'radius' = ONE box: deposit by x-slab + all-reduce(SUM) of the "
                         "accumulator, filter radii split over the GPUs + all-reduce(MAX) of the mask over NCCL")
    ap.add_argument("--no-strong", action="store_true", help="N > 1, partition boxes: skip the one-box strong legs")
    ap.add_argument("--no-share-ics", action="store_true",
                    help="N > 1, end-to-end leg: every rank uploads the whole initial conditions itself")
    ap.add_argument("--strong-steps", type=int, default=5)
    return ap.parse_args()


def workload(args, hii=None):
    hii = hii or args.hii_dim
    dim = (args.dim if args.dim and hii == args.hii_dim else 0) or 3 * hii
    cell = (args.box_len / args.hii_dim) if args.box_len else 1.5
    return hii, dim, cell * hii


def n_radii(hii, box_len, rmax):
    rmin = max(0.620350491, 0.620350491 * box_len / hii)
    rmx = min(rmax, 0.620350491 * box_len)
    return int(np.log(rmx / rmin) / np.log(1.1) + 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def pinned_like(torch, arr):
    t = torch.empty(arr.shape, dtype=torch.float32, pin_memory=True)
    a = t.numpy()
    a[...] = arr
    return t, a


def reference_ics(args, ref, pkg, common, ncpu):
    """Initial conditions for the reference arm, made by the reference itself (outside the timed
    region).  Its IC generator needs ~15 single-threaded FFTs of DIM^3: about half an hour at
    DIM=1536.  The workload's ICs are therefore the reference's own ICs of the half-size box (same
    cell size, HII_DIM/2, DIM/2, BOX_LEN/2) replicated 2x2x2: an exactly periodic field of the
    workload's shape with the same densities, velocities and displacements in cell units."""
    hii, dim, box_len = workload(args)
    t = 2 if (hii >= 256 and hii % 2 == 0 and dim % 2 == 0) else 1
    small = common.make_inputs(hii=hii // t, dim=dim // t, box_len=box_len / t, source=args.source,
                               n_threads=ncpu, R_BUBBLE_MAX=args.r_bubble_max)
    s_ics = pkg.compute_initial_conditions(inputs=small, backend=ref)
    full = common.make_inputs(hii=hii, dim=dim, box_len=box_len, source=args.source, n_threads=ncpu,
                              R_BUBBLE_MAX=args.r_bubble_max)
    if t == 1:
        return full, s_ics, small, s_ics, "the reference's own ComputeInitialConditions"
    ics = pkg.InitialConditions(full)
    for k in ("hires_density", "lowres_density", "lowres_vx", "lowres_vy", "lowres_vz",
              "lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"):
        a = getattr(s_ics, k)
        if a is not None:
            setattr(ics, k, np.ascontiguousarray(np.tile(a, (t, t, t))))
    ics.is_computed = True
    how = (f"the reference's own ComputeInitialConditions at HII_DIM={hii // t} DIM={dim // t} "
           f"BOX_LEN={box_len / t:g}, replicated {t}x{t}x{t} (periodic) to the workload's grid")
    return full, ics, small, s_ics, how


def with_threads(inputs, n):
    """the same inputs with another N_THREADS (the reference sets omp_set_num_threads from it)"""
    return inputs.evolve_input_structs(N_THREADS=int(n))


def cufft_yardstick(torch, lib, hii, box_len, iters=10):
    """SURVEY.md section 8d "library baseline": the transforms of one filter radius done by cuFFT
    (in-place padded r2c / c2r plans through the CUDA toolkit's libcufft, called with ctypes) plus
    the separate passes a cuFFT-based sweep needs around them (window multiply, clip, min/max -- plain
    torch ops), next to the library's own transforms timed alone (b200_fft_probe).  cuFFT is the
    measured baseline here, never the product path."""
    out = {}
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    lib.b200_fft_probe.argtypes = [C.c_int, C.c_int, C.c_double] + [C.POINTER(C.c_double)] * 3
    if lib.b200_fft_probe(hii, iters, float(box_len), C.byref(a), C.byref(b), C.byref(c)) == 0:
        out.update({"own_r2c_ms": a.value, "own_c2r_ms": b.value, "own_c2r_window_clip_minmax_ms": c.value})
    cands = [os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cufft", "lib", "libcufft.so.11"),
             "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so.11", "libcufft.so"]
    cufft = None
    for cand in cands:
        try:
            cufft = C.CDLL(cand)
            break
        except OSError:
            continue
    if cufft is None:
        out["cufft"] = "libcufft not found"
        return out
    nzc = hii // 2 + 1
    box = torch.zeros((hii, hii, 2 * nzc), dtype=torch.float32, device="cuda")
    box[:, :, :hii].normal_()
    win = torch.rand((hii, hii, nzc), dtype=torch.float32, device="cuda")
    kview = torch.view_as_complex(box.view(hii, hii, nzc, 2))
    plans = {}
    for name, typ in (("r2c", 0x2A), ("c2r", 0x2C)):
        h = C.c_int()
        if cufft.cufftPlan3d(C.byref(h), hii, hii, hii, typ) != 0:
            out["cufft"] = "cufftPlan3d failed"
            return out
        cufft.cufftSetStream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        plans[name] = h
    ptr = C.c_void_p(box.data_ptr())

    def timed(fn):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out["cufft_r2c_ms"] = timed(lambda: cufft.cufftExecR2C(plans["r2c"], ptr, ptr))
    out["cufft_c2r_ms"] = timed(lambda: cufft.cufftExecC2R(plans["c2r"], ptr, ptr))

    def sweep():
        kview.mul_(win)                                  # filter_box
        cufft.cufftExecC2R(plans["c2r"], ptr, ptr)        # dft_c2r_cube
        real = box[:, :, :hii]
        torch.aminmax(real)                               # clip_and_get_extrema
        real.clamp_(-1.0, 1e6)
    out["cufft_c2r_window_clip_minmax_ms"] = timed(sweep)
    for h in plans.values():
        cufft.cufftDestroy(h)
    out["note"] = ("per 3-D transform of one HII_DIM^3 box, CUDA events; own = this library's Stockham passes "
                   "(window, clip and min/max fused into them); cufft = in-place cufftExecR2C/C2R + torch "
                   "elementwise passes for the window, clip and extrema")
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) of the SAME workload
    (config.workload) on this host's cores.  One step = ComputePerturbedField + ComputeIonizedBox
    of the full-size box.  A step takes of the order of a minute, so the run is time-boxed
    (BENCH_REF_BUDGET_S, default 240 s of stepping): `steps`/`warmup` echo the request,
    `steps_timed`/`warmup_run` say what was actually run; ms_per_step is the mean of the timed
    steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import common
    pkg = common.pkg
    ref = common.ref_backend()
    ncpu = os.cpu_count() or 1
    hii, dim, box_len = workload(args)
    z = float(args.redshift)
    base = {"impl": "reference", "metric": "coeval cells/sec (perturb+ionize, one redshift)",
            "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"perturb_field+ionize_box z={z} HII_DIM={hii} "
                                   f"DIM={dim} BOX_LEN={box_len:g} {args.source} "
                                   f"n_radii={n_radii(hii, box_len, args.r_bubble_max)}"}}
    if ref is None:
        print(json.dumps({**base, "unavailable": "oracle/_ref/libref21cmfast.so not present"}))
        return
    t_ic = time.perf_counter()
    inputs, ics, small, s_ics, ics_how = reference_ics(args, ref, pkg, common, ncpu)
    t_ic = time.perf_counter() - t_ic

    def step(inp, boxes):
        boxes.inputs = inp
        t0 = time.perf_counter()
        pf = pkg.perturb_field(redshift=z, initial_conditions=boxes, backend=ref)
        t1 = time.perf_counter()
        ib = pkg.compute_ionization_field(perturbed_field=pf, initial_conditions=boxes, backend=ref)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, float(ib.neutral_fraction.mean())

    # thread count: the reference's docs advise few threads (joss-paper/paper.md:261); time one
    # half-size step with 4 threads and with every core, keep the faster setting
    threads, thread_note = ncpu, {}
    if ncpu > 4 and small is not inputs:
        for n in (4, ncpu):
            a, b, _ = step(with_threads(small, n), s_ics)
            thread_note[str(n)] = round(a + b, 3)
        threads = min(thread_note, key=thread_note.get)
        threads = int(threads)
    inputs = with_threads(inputs, threads)
    del s_ics

    budget = float(os.environ.get("BENCH_REF_BUDGET_S", "240"))
    want = args.warmup + args.steps
    done, t_start = [], time.perf_counter()
    while len(done) < want:
        done.append(step(inputs, ics))
        spent = time.perf_counter() - t_start
        if spent + 1.05 * (done[-1][0] + done[-1][1]) > budget:
            break
    # the first step counts as warm-up whenever more than one step fitted into the budget
    n_warm = min(args.warmup, max(0, len(done) - 1), 1 if len(done) < want else args.warmup)
    timed = done[n_warm:]
    ms = 1e3 * float(np.mean([a + b for a, b, _ in timed]))
    value = hii**3 / (ms / 1e3)
    sample = (f"the full workload: HII_DIM={hii} DIM={dim} BOX_LEN={box_len:g}, N_THREADS={threads} "
              f"({ncpu} cores; calibration step seconds by thread count: {thread_note}); {len(timed)} timed "
              f"step(s) after {n_warm} warm-up inside a {budget:g} s budget; perturb "
              f"{np.mean([a for a, _, _ in timed]):.1f} s + ionize {np.mean([b for _, b, _ in timed]):.1f} s; "
              f"FFT back-end = MKL-DFTI shim, single-threaded as in the reference (dft.c), not FFTW; "
              f"ICs = {ics_how} ({t_ic:.0f} s, outside the timed region)")
    print(json.dumps({**base, "value": value, "ms_per_step": ms,
                      "steps_timed": len(timed), "warmup_run": n_warm, "global_xH": timed[-1][2],
                      "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads,
                                       "kind": "reference", "sample": sample},
                      "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": 0}}))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import common
    pkg = common.pkg
    _abi = importlib.import_module("21cmfast_b200._abi")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = pkg.get_backend()
    be.set_table_path(common.table_dir())
    lib = be.lib
    assert lib.b200_set_device(local) == 0
    lib.b200_profile_report.argtypes = [C.c_char_p, C.c_int]

    ncpu = max(1, (os.cpu_count() or 1) // max(1, world))
    hii, dim, box_len = workload(args)
    radius_mode = args.partition in ("radius", "slab") and world > 1  # ONE box over all ranks
    slab_mode = args.partition == "slab" and world > 1
    # N > 1, partition boxes: the ranks work on ONE set of initial conditions, each on its own redshift (a coeval run
    # or lightcone spread over the GPUs) -- what lets the end-to-end leg share the IC upload
    inputs = common.make_inputs(hii=hii, dim=dim, box_len=box_len, source=args.source, seed=1234,
                                n_threads=ncpu, R_BUBBLE_MAX=args.r_bubble_max)
    N, M = hii**3, dim**3
    nrad = n_radii(hii, box_len, args.r_bubble_max)
    os.environ["B200_IC_RNG"] = "device"
    os.environ["B200_SKIP_SCRATCH_OUTPUTS"] = "1"
    ics = pkg.compute_initial_conditions(inputs=inputs, backend=be)
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True)
    z = float(args.redshift)
    z_own = z + (0.25 * rank if (world > 1 and not radius_mode) else 0.0)  # this rank's redshift in the weak legs

    d_ic_part = None

    def stats():
        a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_double()
        lib.b200_last_call_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg (`value`) ----------------
    dev = torch.device("cuda", local)
    names_ic = ["hires_density", "lowres_density", "lowres_vx", "lowres_vy", "lowres_vz",
                "lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
    d_ic = {k: torch.from_numpy(getattr(ics, k)).to(dev) for k in names_ic}
    d_ic_part = {k: v for k, v in d_ic.items() if k != "lowres_density"}
    d_pf = {k: torch.zeros((hii,) * 3, dtype=torch.float32, device=dev) for k in ("density", "velocity_z")}
    d_ib = {k: torch.zeros((hii,) * 3, dtype=torch.float32, device=dev)
            for k in ("neutral_fraction", "z_reion", "kinetic_temperature", "unnormalised_nion")}
    s_ic = _abi.InitialConditionsStruct()
    for k, t in d_ic.items():
        setattr(s_ic, k, C.cast(t.data_ptr(), _abi.c_float_p))
    s_pf = _abi.PerturbedFieldStruct()
    for k, t in d_pf.items():
        setattr(s_pf, k, C.cast(t.data_ptr(), _abi.c_float_p))
    s_ib = _abi.IonizedBoxStruct()
    for k, t in d_ib.items():
        setattr(s_ib, k, C.cast(t.data_ptr(), _abi.c_float_p))
    lib.b200_ComputePerturbedField_device.argtypes = [C.c_float, C.POINTER(_abi.InitialConditionsStruct),
                                                      C.POINTER(_abi.PerturbedFieldStruct)]
    lib.b200_ComputeIonizedBox_device.argtypes = [C.c_float, C.c_float, C.POINTER(_abi.PerturbedFieldStruct),
                                                  C.POINTER(_abi.IonizedBoxStruct)]

    def make_radius_step(ic_part, inp):
        def radius_step():
            """one box on all ranks: slab-parallel deposit + all-reduce(SUM), radius-parallel ionize +
            all-reduce(MAX); host clock around device syncs (library stream + NCCL on torch's stream)"""
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ppf = pkg.perturb_slab_parallel(redshift=z, ics=ic_part, inputs=inp, backend=be)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            out = pkg.ionize_radius_parallel(redshift=z, density=ppf["density"], inputs=inp, backend=be)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            return 1e3 * (t1 - t0), 1e3 * (t2 - t1), 4, float(out["neutral_fraction"].double().sum().item())
        return radius_step

    def make_slab_step(ic_full, inp):
        """one box on x-slabs (library stream only: transposes, halo pull and scalar reductions are peer
        loads / stores inside the library's kernels); device milliseconds of the two library calls"""
        grp = pkg.SlabGroup(inputs=inp, backend=be)
        sl = {k: grp.lowres_slab(v).contiguous() for k, v in ic_full.items() if k.startswith("lowres_v")}
        sl["hires_density"] = grp.hires_slab(ic_full["hires_density"])

        def slab_step():
            ppf = grp.perturb(redshift=z, ics_slab=sl)
            l1, _, _, ms1 = stats()
            out = grp.ionize(redshift=z, density_slab=ppf["density"])
            l2, _, _, ms2 = stats()
            return ms1, ms2, l1 + l2, float(out["neutral_fraction"].double().sum().item())
        return slab_step, grp

    strong_step, slab_group = None, None
    if slab_mode:
        strong_step, slab_group = make_slab_step(d_ic_part, inputs)
    elif radius_mode:
        strong_step = make_radius_step(d_ic_part, inputs)
    strong_xh = [0.0]

    def device_step():
        if radius_mode:
            m1, m2, ln, xs = strong_step()
            strong_xh[0] = xs
            return m1, m2, ln
        d_ib["neutral_fraction"].fill_(1.0)
        d_ib["kinetic_temperature"].zero_()
        torch.cuda.synchronize()
        st = lib.b200_ComputePerturbedField_device(C.c_float(z_own), C.byref(s_ic), C.byref(s_pf))
        assert st == 0, st
        l1, _, _, ms1 = stats()
        st = lib.b200_ComputeIonizedBox_device(C.c_float(z_own), C.c_float(-1.0), C.byref(s_pf), C.byref(s_ib))
        assert st == 0, st
        l2, _, _, ms2 = stats()
        return ms1, ms2, l1 + l2

    for _ in range(args.warmup):
        device_step()
    lib.b200_profile_enable(1)
    clocks = ClockSampler(local)
    clocks.start()
    barrier()
    t_wall = time.perf_counter()
    per, launches = [], 0
    for _ in range(args.steps):
        m1, m2, ln = device_step()
        per.append((m1, m2))
        launches += ln
    barrier()
    t_wall = time.perf_counter() - t_wall
    clk = clocks.stop()
    buf = C.create_string_buffer(1 << 16)
    lib.b200_profile_report(buf, len(buf))
    lib.b200_profile_enable(0)
    prof = {}
    for ln in buf.value.decode().splitlines():
        nm, cnt, tot = ln.rsplit(None, 2)
        nm = nm.split("<")[0]  # template instantiations share one row
        c0, t0 = prof.get(nm, (0, 0.0))
        prof[nm] = (c0 + int(cnt), t0 + float(tot))
    ms_perturb = float(np.mean([p[0] for p in per]))
    ms_ionize = float(np.mean([p[1] for p in per]))
    ms_step = ms_perturb + ms_ionize
    if radius_mode:
        xs = torch.tensor([strong_xh[0]], dtype=torch.float64, device=dev)
        if slab_mode:
            dist.all_reduce(xs)  # every rank holds a slab; the radius partition leaves the whole box everywhere
        xh_dev = float(xs.item()) / N
    else:
        xh_dev = float(d_ib["neutral_fraction"].mean().item())

    # ---------------- end-to-end leg through the C-ABI with pinned host buffers ----------------
    e2e = None
    if not args.no_e2e and not radius_mode:
        os.environ["B200_ICS_CACHE"] = "0"  # (the library default) every step uploads its initial conditions
        keep = []
        h_ics = pkg.InitialConditions(inputs)
        for k in names_ic:
            t, a = pinned_like(torch, getattr(ics, k))
            keep.append(t)
            setattr(h_ics, k, a)
        h_pf = pkg.PerturbedField(inputs, z_own)
        h_ib = pkg.IonizedBox(inputs, z_own)
        h_prev = pkg.IonizedBox(inputs, -1.0)
        for obj, ks in ((h_pf, ("density", "velocity_z")),
                        (h_ib, ("neutral_fraction", "ionisation_rate_G12", "mean_free_path", "z_reion",
                                "kinetic_temperature", "unnormalised_nion")), (h_prev, ("z_reion",))):
            for k in ks:
                t, a = pinned_like(torch, np.zeros((hii,) * 3, np.float32))
                keep.append(t)
                setattr(obj, k, a)
        ppf, ts, hb = pkg.PerturbedField(inputs, -1.0), pkg.outputs.TsBox.dummy(inputs), pkg.outputs.HaloBox.dummy(inputs)

        e2e_split = []  # seconds inside ComputePerturbedField, per step

        def host_step():
            h_ib.neutral_fraction[...] = 1.0
            h_ib.kinetic_temperature[...] = 0.0
            t0 = time.perf_counter()
            st = lib.ComputePerturbedField(C.c_float(z_own), C.byref(h_ics.cstruct), C.byref(h_pf.cstruct))
            assert st == 0, st
            e2e_split.append(time.perf_counter() - t0)
            _, hb1, db1, _ = stats()
            st = lib.ComputeIonizedBox(C.c_float(z_own), C.c_float(-1.0), C.byref(h_pf.cstruct), C.byref(ppf.cstruct),
                                       C.byref(h_prev.cstruct), C.byref(ts.cstruct), C.byref(hb.cstruct),
                                       C.byref(h_ics.cstruct), C.byref(h_ib.cstruct))
            assert st == 0, st
            _, hb2, db2, _ = stats()
            return time.perf_counter() - t0, hb1 + hb2, db1 + db2

        # N > 1: the ranks' boxes share their initial conditions, so each rank uploads 1 / N of them over its own
        # PCIe link and the shares travel to the peers over NVLink (b200_ics_share); the C-ABI calls stay the same
        share_grp = None
        if world > 1 and not args.no_share_ics:
            share_grp = pkg.SlabGroup(inputs=inputs, backend=be, heap_bytes=pkg.SlabGroup.ics_heap_bytes(inputs))
            share_grp.share_ics(True)
        for _ in range(max(1, args.warmup - 1)):
            host_step()
        barrier()
        tt, h2d_b, d2h_b = [], 0, 0
        for _ in range(args.steps):
            t, hb_, db_ = host_step()
            tt.append(t)
            h2d_b, d2h_b = hb_, db_
        barrier()
        if share_grp is not None:
            share_grp.close()
        e2e_s = float(np.mean(tt))
        xh_host = float(h_ib.neutral_fraction.mean())
        assert abs(xh_host - xh_dev) < 1e-6, (xh_host, xh_dev)
        # what the link itself does: one 2 GiB pinned copy each way on torch's stream (the e2e floor is
        # h2d_bytes / this rate when the two directions overlap perfectly)
        big = torch.empty(1 << 29, dtype=torch.float32, pin_memory=True)
        dbig = torch.empty(1 << 29, dtype=torch.float32, device=dev)
        link = {}
        for name, (dst, src) in (("h2d", (dbig, big)), ("d2h", (big, dbig))):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            link[name + "_GBps"] = big.numel() * 4 / (time.perf_counter() - t0) / 1e9
        del big, dbig
        e2e = (e2e_s, h2d_b, d2h_b, float(np.mean(e2e_split[-args.steps:])), link)
        os.environ.pop("B200_ICS_CACHE")

    # ---------------- ONE box over all GPUs (strong scaling), reported beside the replica value ----------------
    strong = None
    if world > 1 and not radius_mode and not args.no_strong:
        strong = {}
        del d_ic_part
        cin = inputs  # every rank holds the same initial conditions (seed 1234)
        d_c = {k: v for k, v in d_ic.items() if k != "lowres_density"}
        be.state.init(cin, broadcast_inputs=True, ps=True, sigma=True, heat=True)
        xh0 = torch.tensor([xh_dev], dtype=torch.float64, device=dev)
        dist.broadcast(xh0, 0)  # the single-GPU answer for the same box at the same redshift (rank 0's)
        for name in ("slab", "radius"):
            try:
                if name == "slab":
                    step, grp = make_slab_step(d_c, cin)
                else:
                    step, grp = make_radius_step(d_c, cin), None
                for _ in range(2):
                    step()
                barrier()
                t0 = time.perf_counter()
                rec = [step() for _ in range(args.strong_steps)]
                barrier()
                wall = (time.perf_counter() - t0) / args.strong_steps
                v = torch.tensor([np.mean([r[0] for r in rec]), np.mean([r[1] for r in rec]), wall, rec[-1][3]],
                                 dtype=torch.float64, device=dev)
                vmax = v.clone()
                dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
                xs = v[3:4].clone()
                if name == "slab":
                    dist.all_reduce(xs)
                ms_s = float(vmax[0] + vmax[1])
                strong[name] = {"ms_per_step": ms_s, "ms_perturb": float(vmax[0]), "ms_ionize": float(vmax[1]),
                                "wall_ms_per_step": 1e3 * float(vmax[2]), "value": N / (ms_s * 1e-3), "unit": "cells/s",
                                "scaling": "strong", "steps": args.strong_steps,
                                "global_xH": float(xs.item()) / N,
                                "matches_single_gpu_xH": bool(abs(float(xs.item()) / N - float(xh0.item())) < 1e-7),
                                "speedup_vs_one_gpu": ms_step / ms_s}
                if grp is not None:
                    grp.close()
            except Exception as e:  # a strong leg must never cost the replica line
                strong[name] = {"error": repr(e)[:300]}
                break

    # ---------------- reduce over ranks (max time) ----------------
    vals = torch.tensor([ms_step, ms_perturb, ms_ionize, e2e[0] if e2e else 0.0, t_wall], device=dev,
                        dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_step, ms_perturb, ms_ionize, e2e_s, t_wall = [float(x) for x in vals.tolist()]

    yardstick = None
    if rank == 0 and world == 1 and not args.no_yardstick:
        torch.cuda.synchronize()
        yardstick = cufft_yardstick(torch, lib, hii, box_len)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = common.ref_backend()
        if ref is not None:
            rh, rd, rb = workload(args, args.ref_hii_dim)
            rin = common.make_inputs(hii=rh, dim=rd, box_len=rb, source=args.source, n_threads=os.cpu_count() or 1,
                                     R_BUBBLE_MAX=args.r_bubble_max)
            os.environ["B200_IC_RNG"] = "device"
            rics = pkg.compute_initial_conditions(inputs=rin, backend=be)
            t0 = time.perf_counter()
            rpf = pkg.perturb_field(redshift=z, initial_conditions=rics, backend=ref)
            t1 = time.perf_counter()
            pkg.compute_ionization_field(perturbed_field=rpf, initial_conditions=rics, backend=ref)
            t2 = time.perf_counter()
            cpu_baseline = {
                "value": rh**3 / (t2 - t0), "unit": "cells/s", "cores": os.cpu_count() or 1, "kind": "reference",
                "sample": (f"oracle/_ref (reference C + FFTW/GSL shims, MKL-DFTI FFT single-threaded as in dft.c) "
                           f"HII_DIM={rh} DIM={rd} BOX_LEN={rb:g} z={z}, one run: perturb {t1 - t0:.2f}s "
                           f"ionize {t2 - t1:.2f}s, N_THREADS={os.cpu_count()}")}

    if rank == 0:
        peak, peak_kind = measured_peak()
        pitch = ((hii // 2 + 1) + 7) // 8 * 8
        Nk = hii * hii * (hii // 2 + 1)  # algorithmic = unpadded modes (the row pitch is an internal choice)
        alg = {  # algorithmic bytes per launch
            "fft_strided_pow2_kernel": 16 * Nk, "fft_strided_kernel": 16 * Nk,
            "fft_c2r_z_pow2_kernel": 8 * Nk + 4 * N, "fft_c2r_z_kernel": 8 * Nk + 4 * N,
            "fft_r2c_z_pow2_kernel": 8 * Nk + 4 * N, "fft_r2c_z_kernel": 8 * Nk + 4 * N,
            "fcoll_sum_kernel": 4 * N, "ionise_delta_kernel": 4 * N, "ionise_kernel": 4 * N,
            "window_expand_kernel": 4 * (hii // 2 + 1) ** 2 * pitch,
            "move_cic_grouped_kernel": 4 * M + 24 * N + 8 * N, "move_cic_kernel": 4 * M + 24 * N + 8 * N,
            "acc_to_delta_kernel": 12 * N, "finalize_kernel": 17 * N, "fill_kernel": 4 * N}
        dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else None
        roofline = None
        if dom:
            nm, (cnt, tot) = dom
            avg_ms = tot / cnt
            ach = alg.get(nm, 0) / (avg_ms * 1e-3) / 1e9
            traffic = None
            tp = ROOT / "profiles" / "traffic.json"
            if tp.exists():
                traffic = json.loads(tp.read_text()).get(f"{nm}@{hii}")
            roofline = {"bound": "hbm", "kernel": nm, "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": traffic, "peak_kind": f"of {peak_kind}",
                        "launches": cnt, "avg_launch_ms": avg_ms,
                        "share_of_kernel_time": tot / sum(v[1] for v in prof.values())}
        step_bytes = (40 + 24 * nrad) * N + (4 * M + 72 * N)
        out = {
            "metric": "coeval cells/sec (perturb+ionize, one redshift)",
            "value": (1 if radius_mode else world) * N / (ms_step * 1e-3),
            "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if radius_mode else "weak",
            "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"perturb_field+ionize_box z={z} HII_DIM={hii} DIM={dim} BOX_LEN={box_len:g} "
                                   f"{args.source} n_radii={nrad}",
                       "parallelism": (f"one box over {world} GPUs on x-slabs: slab-decomposed FFTs (transposes = peer "
                                       f"stores over NVLink), slab deposit + halo pull") if slab_mode else
                                      (f"one box over {world} GPUs: x-slab deposit + all-reduce(SUM), radii split + "
                                       f"all-reduce(MAX) of the mask") if radius_mode else
                                      (f"{world} coeval boxes at {world} redshifts (z + 0.25 rank) on shared initial "
                                       f"conditions, one per GPU" if world > 1 else "1 coeval box"),
                       "l2": f"inputs larger than L2 (every pass streams a {4 * N / 1e6:.0f} MB box; L2 is 126 MB)",
                       "ms_perturb": ms_perturb, "ms_ionize": ms_ionize, "global_xH": xh_dev,
                       "wall_ms_per_step": 1e3 * t_wall / args.steps},
            "clocks": clk, "gpu_launches": launches,
            "roofline": roofline,
            # one box over `world` GPUs is measured against world x the per-GPU peak
            "step_roofline": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                              "peak": peak * (world if radius_mode else 1), "unit": "GB/s",
                              "frac": step_bytes / (ms_step * 1e-3) / 1e9 / (peak * (world if radius_mode else 1))},
            "kernel_profile_ms_per_step": {k: v[1] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
            "cpu_baseline": cpu_baseline,
            "cufft_yardstick": yardstick,
            "strong": strong,
        }
        if e2e:
            out["e2e"] = {"value": world * N / e2e_s, "unit": "cells/s", "h2d_bytes_per_step": int(e2e[1]),
                          "d2h_bytes_per_step": int(e2e[2]), "ms_per_step": 1e3 * e2e_s,
                          "ms_perturb_call": 1e3 * e2e[3], "ms_ionize_call": 1e3 * (e2e_s - e2e[3]),
                          "ics_upload": ("shared: h2d bytes are this rank's 1/N share of the initial conditions, the "
                                         "rest arrives from the peers over NVLink" if (world > 1 and not args.no_share_ics)
                                         else "whole initial conditions per call"),
                          "pinned_copy_yardstick": e2e[4]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
