"""ONE coeval box across the GPUs of a node: slab-parallel particle deposit
and radius-parallel ionisation.

Deposit: ``move_grid_masses`` (reference ``map_mass.c:146-208``) is a sum over particles; rank ``r``
deposits the particles of its x-slab into a fixed-point 64-bit accumulator and ONE
``all_reduce(SUM)`` over int64 merges them -- integer addition, so the merged grid is bit-identical
to the single-GPU deposit.  In a host-fed deployment every rank uploads only its slab of
``hires_density`` (its own PCIe link).

Ionisation:

Given the k-space density the filter radii of ``find_HII_bubbles`` (reference
``IonisationBox.c:1531-1630``) are independent: every radius needs only its own filtered grid,
collapse-fraction table, grid sum and flags, and the flags combine by OR.  Rank ``r`` of ``P``
therefore runs the radii ``k = r (mod P)`` (phase 0), the byte masks are combined with ONE
``all_reduce(MAX)`` over NCCL/NVLink (134 MB at 512^3), and every rank runs the last radius, which
assigns the partial ionisations of the never-flagged cells, on the merged mask (phase 1).

PyTorch is plumbing only: it owns the device tensors and the process group.  The computation is
``b200_ComputeIonizedBox_device_part`` of the C-ABI library (``include/py21cmfast_b200.h``).
The same code runs on CPU tensors against the host-emulation library (gloo), which is how the
test-suite covers the N > 1 logic without a GPU.
"""
from __future__ import annotations

import ctypes as C

from . import _abi
from ._lib import BackendError


def _ptr(t):
    return C.cast(t.data_ptr(), _abi.c_float_p)


def ionize_radius_parallel(*, redshift: float, density, inputs, backend, group=None,
                           want_nion: bool = True):
    """Ionise one box cooperatively.

    density : float32 tensor ``(HII_DIM, HII_DIM, HII_D_PARA)`` resident where the backend computes
              (CUDA for the product library); identical on every rank.
    Returns ``dict(neutral_fraction, z_reion, kinetic_temperature, unnormalised_nion, mean_f_coll)``
    -- complete on every rank.
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    be = backend
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True)
    lib = be.lib
    lib.b200_ComputeIonizedBox_device_part.argtypes = [
        C.c_float, C.c_float, C.POINTER(_abi.PerturbedFieldStruct), C.POINTER(_abi.IonizedBoxStruct),
        C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.b200_ComputeIonizedBox_device_part.restype = C.c_int

    dev, shape = density.device, tuple(density.shape)
    out = {
        "neutral_fraction": torch.ones(shape, dtype=torch.float32, device=dev),
        "z_reion": torch.zeros(shape, dtype=torch.float32, device=dev),
        "kinetic_temperature": torch.zeros(shape, dtype=torch.float32, device=dev),
    }
    if want_nion:
        out["unnormalised_nion"] = torch.zeros(shape, dtype=torch.float32, device=dev)
    mask = torch.zeros(density.numel(), dtype=torch.uint8, device=dev)
    s_pf = _abi.PerturbedFieldStruct()
    s_pf.density = _ptr(density)
    s_ib = _abi.IonizedBoxStruct()
    for k, t in out.items():
        setattr(s_ib, k, _ptr(t))
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)  # the library runs on its own stream

    def call(phase):
        st = lib.b200_ComputeIonizedBox_device_part(
            C.c_float(redshift), C.c_float(-1.0), C.byref(s_pf), C.byref(s_ib),
            C.c_void_p(mask.data_ptr()), rank, world, phase)
        if st != 0:
            raise BackendError(st, f"b200_ComputeIonizedBox_device_part(phase={phase})")

    call(0)
    if world > 1:
        dist.all_reduce(mask, op=dist.ReduceOp.MAX, group=group)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
    call(1)
    out["mean_f_coll"] = float(s_ib.mean_f_coll)
    return out


def perturb_slab_parallel(*, redshift: float, ics: dict, inputs, backend, group=None):
    """Perturb one box cooperatively.

    ics : dict of float32 tensors resident where the backend computes -- ``hires_density``
          ``(DIM, DIM, D_PARA)`` (only the planes of the rank's own x-slab and the one below it are
          read in phase 0), ``lowres_vx/vy/vz`` and, for 2LPT, ``lowres_vx_2LPT/...``.
    Returns ``dict(density, velocity_z)`` -- complete on every rank.
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    be = backend
    be.state.init(inputs, broadcast_inputs=True)
    lib = be.lib
    lib.b200_ComputePerturbedField_device_part.argtypes = [
        C.c_float, C.POINTER(_abi.InitialConditionsStruct), C.POINTER(_abi.PerturbedFieldStruct),
        C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.b200_ComputePerturbedField_device_part.restype = C.c_int

    so = inputs.simulation_options
    hii = so.HII_DIM
    shape = tuple(ics["lowres_vx"].shape)
    dev = ics["lowres_vx"].device
    out = {"density": torch.zeros(shape, dtype=torch.float32, device=dev),
           "velocity_z": torch.zeros(shape, dtype=torch.float32, device=dev)}
    acc = torch.zeros(hii * shape[1] * shape[2], dtype=torch.int64, device=dev)
    s_ic = _abi.InitialConditionsStruct()
    for k, t in ics.items():
        setattr(s_ic, k, _ptr(t))
    s_pf = _abi.PerturbedFieldStruct()
    for k, t in out.items():
        setattr(s_pf, k, _ptr(t))
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)

    def call(phase):
        st = lib.b200_ComputePerturbedField_device_part(
            C.c_float(redshift), C.byref(s_ic), C.byref(s_pf), C.c_void_p(acc.data_ptr()), rank, world, phase)
        if st != 0:
            raise BackendError(st, f"b200_ComputePerturbedField_device_part(phase={phase})")

    call(0)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
    call(1)
    return out
