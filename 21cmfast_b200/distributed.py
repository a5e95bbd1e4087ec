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
        _agree(st, f"b200_ComputeIonizedBox_device_part(phase={phase})", group)

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
        _agree(st, f"b200_ComputePerturbedField_device_part(phase={phase})", group)

    call(0)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
    call(1)
    return out


# ---------------------------------------------------------------------------------------------
# Slab decomposition: ONE box, every stage on x-slabs, transposes / halos over peer memory
# ---------------------------------------------------------------------------------------------
def _agree(status: int, where: str, group=None):
    """Raise on EVERY rank when any rank failed (a lone raise would leave the others in a collective)."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([int(status)], dtype=torch.int32)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        status = int(t.item())
    if status != 0:
        raise BackendError(status, where)


class SlabGroup:
    """The ranks of ``torch.distributed`` as one slab-decomposed box (``include/py21cmfast_b200.h``,
    "ONE box over the GPUs of a node").

    Construction allocates the library's symmetric heap on this rank's device, exchanges the 64-byte
    peer-memory handles through ``torch.distributed`` (the only use of it on this path) and maps the
    peers.  Afterwards ``perturb`` / ``ionize`` are plain library calls: the all-to-all transposes of
    the FFTs, the halo exchange of the deposit and the per-radius scalar reductions all happen inside
    the library's kernels over NVLink.
    """

    def __init__(self, *, inputs, backend, group=None, heap_bytes: int | None = None, ics: bool = False):
        import torch.distributed as dist
        self.backend, self.inputs, self.group = backend, inputs, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        so = inputs.simulation_options
        hii, hz = so.HII_DIM, so.HII_D_PARA
        if hii % self.world:
            raise ValueError(f"HII_DIM={hii} must be a multiple of the number of ranks ({self.world})")
        self.nxl = hii // self.world
        self.x0 = self.rank * self.nxl
        self.F = so.dim // hii
        if heap_bytes is None:
            pitch = ((hz // 2 + 1) + 7) // 8 * 8
            halo = min(self.nxl, 24)
            heap_bytes = (2 * 8 * self.nxl * hii * pitch + 8 * (self.nxl + 2 * halo) * hii * hz + (16 << 20))
            if ics:  # slab ICs transform the hi-res box: two receive buffers and the two gathered Hermitian planes
                dim, dz = so.dim, so.D_PARA
                dpitch = ((dz // 2 + 1) + 7) // 8 * 8
                heap_bytes = max(heap_bytes, 2 * 8 * (dim // self.world) * dim * dpitch + 16 * dim * dim + (16 << 20))
        lib = backend.lib
        lib.b200_dist_init.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_void_p]
        lib.b200_dist_connect.argtypes = [C.c_void_p]
        lib.b200_ComputePerturbedField_slab.argtypes = [
            C.c_float, C.POINTER(_abi.InitialConditionsStruct), C.POINTER(_abi.PerturbedFieldStruct)]
        lib.b200_ComputeIonizedBox_slab.argtypes = [
            C.c_float, C.c_float, C.POINTER(_abi.PerturbedFieldStruct), C.POINTER(_abi.IonizedBoxStruct)]
        lib.b200_ComputeInitialConditions_slab.argtypes = [C.c_ulonglong, C.POINTER(_abi.InitialConditionsStruct)]
        handle = C.create_string_buffer(64)
        st = lib.b200_dist_init(self.rank, self.world, C.c_ulonglong(heap_bytes), handle)
        _agree(st, "b200_dist_init", group)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, handle.raw, group=group)
        else:
            handles = [handle.raw]
        blob = C.create_string_buffer(b"".join(handles), 64 * self.world)
        st = lib.b200_dist_connect(blob)
        _agree(st, "b200_dist_connect", group)
        self.open = True

    def close(self):
        if self.open:
            self.share_ics(False)
            self.backend.lib.b200_dist_shutdown()
            self.open = False

    def share_ics(self, enable: bool = True):
        """Redshift-parallel runs on shared initial conditions (``b200_ics_share``): while enabled, the ranks call
        ``perturb_field`` / ``ComputePerturbedField`` together, each with its own redshift but the same IC arrays in
        host memory; every rank uploads 1 / world of them and the shares travel to the peers over NVLink.  The
        group's heap must hold the ICs (``SlabGroup.ics_heap_bytes(inputs)``)."""
        fn = self.backend.lib.b200_ics_share
        fn.argtypes, fn.restype = [C.c_int], None
        fn(1 if enable else 0)

    @staticmethod
    def ics_heap_bytes(inputs) -> int:
        so = inputs.simulation_options
        n_lo, n_hi = so.HII_DIM * so.HII_DIM * so.HII_D_PARA, so.dim * so.dim * so.D_PARA
        return 4 * (n_hi + 6 * n_lo) + (64 << 20)

    # -- slab views of whole-box arrays (host-side helpers for tests / feeders) ------------------
    def lowres_slab(self, a):
        return a[self.x0:self.x0 + self.nxl]

    def hires_slab(self, a):
        """the F * nxl hi-res planes the rank's velocity cells read: from F x0 - F // 2, periodic"""
        import numpy as np
        n = a.shape[0]
        idx = (np.arange(self.F * self.nxl) + self.F * self.x0 - self.F // 2) % n
        if isinstance(a, np.ndarray):
            return np.ascontiguousarray(a[idx])
        import torch
        return a[torch.as_tensor(idx, device=a.device)].contiguous()

    def initial_conditions(self, *, device=None):
        """The initial conditions of the box generated slab by slab (``b200_ComputeInitialConditions_slab``): the
        hi-res box never exists on one GPU.  Returns this rank's x-slabs as tensors on ``device``:
        ``hires_density`` (planes ``DIM / world * rank ...``) and the low-res density / velocity boxes
        (planes ``x0 ... x0 + nxl``).  The group must have been built with ``ics=True`` (heap size)."""
        import torch
        be, inputs = self.backend, self.inputs
        so, mo = inputs.simulation_options, inputs.matter_options
        be.state.init(inputs, broadcast_inputs=True, ps=True)
        dev = torch.device(device) if device is not None else torch.device("cpu")
        lo = (self.nxl, so.HII_DIM, so.HII_D_PARA)
        hi = (so.dim // self.world, so.dim, so.D_PARA)
        names = ["lowres_density", "lowres_vx", "lowres_vy", "lowres_vz"]
        if mo.PERTURB_ALGORITHM == "2LPT":
            names += ["lowres_vx_2LPT", "lowres_vy_2LPT", "lowres_vz_2LPT"]
        out = {k: torch.zeros(lo, dtype=torch.float32, device=dev) for k in names}
        out["hires_density"] = torch.zeros(hi, dtype=torch.float32, device=dev)
        s_ic = _abi.InitialConditionsStruct()
        for k, t in out.items():
            setattr(s_ic, k, _ptr(t))
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        st = be.lib.b200_ComputeInitialConditions_slab(C.c_ulonglong(inputs.random_seed), C.byref(s_ic))
        _agree(st, "b200_ComputeInitialConditions_slab", self.group)
        return out

    def shift_hires(self, hires_natural):
        """The hi-res planes ``perturb`` reads start half a low-res cell below the rank's slab (``hires_slab``); the
        slab-decomposed ICs deliver the planes ``F x0 ... F (x0 + nxl)``.  Fetches the last ``F // 2`` planes of the
        previous rank (periodic) with one small all-gather and returns the shifted slab."""
        import torch
        import torch.distributed as dist
        h = self.F // 2
        if h == 0:
            return hires_natural
        tail = hires_natural[-h:].contiguous()
        if self.world > 1:
            tails = [torch.empty_like(tail) for _ in range(self.world)]
            dist.all_gather(tails, tail, group=self.group)
            prev = tails[(self.rank - 1) % self.world]
        else:
            prev = tail
        return torch.cat([prev, hires_natural[:-h]]).contiguous()

    def perturb(self, *, redshift: float, ics_slab: dict):
        """ics_slab: device tensors of this rank's slabs (``hires_density`` from ``hires_slab``, the low-res
        velocity boxes from ``lowres_slab``; ``PERTURB_ALGORITHM='LINEAR'`` reads ``lowres_density`` instead).
        Returns ``dict(density, velocity_z)`` slabs."""
        import torch
        be = self.backend
        be.state.init(self.inputs, broadcast_inputs=True)
        ref = ics_slab["lowres_vx"]
        dev, shape = ref.device, tuple(ref.shape)
        out = {"density": torch.zeros(shape, dtype=torch.float32, device=dev),
               "velocity_z": torch.zeros(shape, dtype=torch.float32, device=dev)}
        s_ic = _abi.InitialConditionsStruct()
        for k, t in ics_slab.items():
            assert t.is_contiguous()
            setattr(s_ic, k, _ptr(t))
        s_pf = _abi.PerturbedFieldStruct()
        for k, t in out.items():
            setattr(s_pf, k, _ptr(t))
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        st = be.lib.b200_ComputePerturbedField_slab(C.c_float(redshift), C.byref(s_ic), C.byref(s_pf))
        _agree(st, "b200_ComputePerturbedField_slab", self.group)
        return out

    def ionize(self, *, redshift: float, density_slab, want_nion: bool = True):
        """density_slab: this rank's x-slab of the perturbed density (device tensor).  Returns the slabs of
        ``neutral_fraction``, ``z_reion``, ``kinetic_temperature`` (, ``unnormalised_nion``) and ``mean_f_coll``."""
        import torch
        be = self.backend
        be.state.init(self.inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True)
        dev, shape = density_slab.device, tuple(density_slab.shape)
        out = {"neutral_fraction": torch.ones(shape, dtype=torch.float32, device=dev),
               "z_reion": torch.zeros(shape, dtype=torch.float32, device=dev),
               "kinetic_temperature": torch.zeros(shape, dtype=torch.float32, device=dev)}
        if want_nion:
            out["unnormalised_nion"] = torch.zeros(shape, dtype=torch.float32, device=dev)
        s_pf = _abi.PerturbedFieldStruct()
        s_pf.density = _ptr(density_slab)
        s_ib = _abi.IonizedBoxStruct()
        for k, t in out.items():
            setattr(s_ib, k, _ptr(t))
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        st = be.lib.b200_ComputeIonizedBox_slab(C.c_float(redshift), C.c_float(-1.0), C.byref(s_pf), C.byref(s_ib))
        _agree(st, "b200_ComputeIonizedBox_slab", self.group)
        out["mean_f_coll"] = float(s_ib.mean_f_coll)
        return out
