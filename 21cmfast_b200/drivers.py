"""Single-field drivers mirroring ``py21cmfast.drivers.single_field`` for the hot path.

``compute_initial_conditions`` (single_field.py:38-113), ``perturb_field`` (:116-156),
``compute_ionization_field`` (:714-840), ``brightness_temperature`` and a minimal ``run_coeval``
(coeval.py:521-697: independent redshifts, or the scrolled evolution over node redshifts that
``RECOMB_MODEL`` needs).  Each call initialises the backend's global
state exactly as ``@init_c_state`` does in the reference and then calls the C-ABI entry point
with numpy-owned host buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Backend, BackendError, get_backend
from .inputs import InputParameters
from .outputs import (BrightnessTemp, HaloBox, InitialConditions, IonizedBox, PerturbedField,
                      TsBox)


def _check(status, where):
    if status != 0:
        raise BackendError(status, where)


def compute_initial_conditions(*, inputs: InputParameters, backend: Backend | None = None,
                               initial_density=None) -> InitialConditions:
    be = backend or get_backend()
    be.state.init(inputs, broadcast_inputs=True, ps=True)
    ics = InitialConditions.new(inputs)
    if initial_density is not None:  # "given hires_density" branch, InitialConditions.c:636-663
        ics.hires_density[...] = initial_density
    _check(be.lib.ComputeInitialConditions(C.c_ulonglong(inputs.random_seed), C.byref(ics.cstruct)),
           "ComputeInitialConditions")
    ics.is_computed = True
    return ics


def perturb_field(*, redshift: float, initial_conditions: InitialConditions,
                  backend: Backend | None = None) -> PerturbedField:
    be = backend or get_backend()
    inputs = initial_conditions.inputs
    be.state.init(inputs, broadcast_inputs=True)
    pf = PerturbedField.new(inputs, redshift)
    _check(be.lib.ComputePerturbedField(C.c_float(redshift), C.byref(initial_conditions.cstruct),
                                        C.byref(pf.cstruct)), "ComputePerturbedField")
    pf.is_computed = True
    return pf


class _HaloCatalogStruct(C.Structure):
    """``HaloCatalog.dummy()``: no sampled halos (``_outputstructs_wrapper.h:30-45``); L-INTEGRAL never reads it."""
    _fields_ = [("n_halos", C.c_ulonglong), ("buffer_size", C.c_ulonglong), ("halo_masses", C.c_void_p),
                ("halo_coords", C.c_void_p), ("star_rng", C.c_void_p), ("sfr_rng", C.c_void_p), ("xray_rng", C.c_void_p)]


def compute_halobox(*, redshift: float, initial_conditions: InitialConditions,
                    previous_spin_temp: TsBox | None = None, previous_ionize_box: IonizedBox | None = None,
                    backend: Backend | None = None) -> HaloBox:
    """``compute_halo_grid`` (single_field.py:297-380) for ``SOURCE_MODEL='L-INTEGRAL'``: the photon-output and
    star-formation grids integrated over the conditional mass function of every Lagrangian cell and moved to
    Eulerian space with the perturbed field's displacements.  Sampled halo catalogues are out of scope."""
    be = backend or get_backend()
    inputs = initial_conditions.inputs
    if inputs.matter_options.SOURCE_MODEL != "L-INTEGRAL":
        raise NotImplementedError("only SOURCE_MODEL='L-INTEGRAL' builds a HaloBox here (no halo sampler)")
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True)
    hb = HaloBox.new(inputs, redshift)
    cat = _HaloCatalogStruct()
    ts = previous_spin_temp or TsBox.dummy(inputs)
    prev = previous_ionize_box or IonizedBox(inputs, -1.0)
    fn = be.lib.ComputeHaloBox
    fn.restype = C.c_int
    _check(fn(C.c_double(redshift), C.byref(initial_conditions.cstruct), C.byref(cat), C.byref(ts.cstruct),
              C.byref(prev.cstruct), C.byref(hb.cstruct)), "ComputeHaloBox")
    hb.pull_scalars()
    hb.is_computed = True
    return hb


def compute_halo_grid(*, redshift: float, initial_conditions: InitialConditions, inputs: InputParameters | None = None,
                      halo_catalog=None, previous_spin_temp: TsBox | None = None,
                      previous_ionize_box: IonizedBox | None = None, backend: Backend | None = None) -> HaloBox:
    """The reference's name and keywords for ``compute_halobox`` (``drivers/single_field.py:297-380``).  A sampled
    ``halo_catalog`` is outside the scoped path: only the integrated grids of ``SOURCE_MODEL='L-INTEGRAL'`` are built."""
    if halo_catalog is not None:
        raise NotImplementedError("sampled halo catalogues are outside the scoped path (SURVEY.md section 8)")
    if inputs is not None and inputs != initial_conditions.inputs:
        raise ValueError("inputs differ from the ones the initial conditions were made with")
    return compute_halobox(redshift=redshift, initial_conditions=initial_conditions, previous_spin_temp=previous_spin_temp,
                           previous_ionize_box=previous_ionize_box, backend=backend)


def compute_ionization_field(*, perturbed_field: PerturbedField,
                             initial_conditions: InitialConditions,
                             previous_perturbed_field: PerturbedField | None = None,
                             previous_ionized_box: IonizedBox | None = None,
                             spin_temp: TsBox | None = None,
                             halobox: HaloBox | None = None,
                             backend: Backend | None = None) -> IonizedBox:
    be = backend or get_backend()
    inputs = perturbed_field.inputs
    ao = inputs.astro_options
    lagrangian = inputs.matter_options.lagrangian_source_grid
    if ao.USE_MINI_HALOS or (lagrangian and inputs.matter_options.SOURCE_MODEL != "L-INTEGRAL"):
        raise NotImplementedError(
            "mini-halos and the halo-sampler source models are outside the scoped IonizeBox path (SURVEY.md section 8)")
    if lagrangian and halobox is None:  # single_field.py:796-801
        raise ValueError("SOURCE_MODEL requires a halobox, but none was provided")
    if ao.USE_TS_FLUCT and spin_temp is None:  # single_field.py:803-808
        raise ValueError("You have USE_TS_FLUCT=True, but have not provided a spin_temp!")
    redshift = perturbed_field.redshift
    # previous-snapshot rules of the reference (single_field.py:773-791)
    if redshift >= inputs.simulation_options.Z_HEAT_MAX:
        previous_ionized_box = IonizedBox.initial(inputs)
        previous_perturbed_field = PerturbedField.initial(inputs)
    if inputs.evolution_required:
        if previous_ionized_box is None:
            raise ValueError("You need to provide a previous ionized box when redshift < Z_HEAT_MAX.")
        if previous_perturbed_field is None:
            raise ValueError("You need to provide a previous perturbed field when redshift < Z_HEAT_MAX.")
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True, heat=True, recomb=True)
    prev_pf = previous_perturbed_field or PerturbedField.initial(inputs)
    prev_ion = previous_ionized_box or IonizedBox.initial(inputs)
    ts, hb = (spin_temp if ao.USE_TS_FLUCT else TsBox.dummy(inputs)), (halobox if lagrangian else HaloBox.dummy(inputs))
    box = IonizedBox.new(inputs, redshift)
    _check(be.lib.ComputeIonizedBox(
        C.c_float(redshift), C.c_float(prev_pf.redshift), C.byref(perturbed_field.cstruct),
        C.byref(prev_pf.cstruct), C.byref(prev_ion.cstruct), C.byref(ts.cstruct),
        C.byref(hb.cstruct), C.byref(initial_conditions.cstruct), C.byref(box.cstruct)),
        "ComputeIonizedBox")
    box.pull_scalars()
    box.is_computed = True
    return box


def brightness_temperature(*, ionized_box: IonizedBox, perturbed_field: PerturbedField,
                           spin_temp: TsBox | None = None,
                           backend: Backend | None = None) -> BrightnessTemp:
    be = backend or get_backend()
    inputs = ionized_box.inputs
    if inputs.astro_options.USE_TS_FLUCT and spin_temp is None:  # single_field.py brightness_temperature
        raise ValueError("You have USE_TS_FLUCT=True, but have not provided a spin_temp!")
    be.state.init(inputs, broadcast_inputs=True)
    bt = BrightnessTemp.new(inputs, ionized_box.redshift)
    ts = spin_temp if inputs.astro_options.USE_TS_FLUCT else TsBox.dummy(inputs)
    _check(be.lib.ComputeBrightnessTemp(
        C.c_float(ionized_box.redshift), C.byref(ts.cstruct), C.byref(ionized_box.cstruct),
        C.byref(perturbed_field.cstruct), C.byref(bt.cstruct)), "ComputeBrightnessTemp")
    bt.is_computed = True
    return bt


def get_logspaced_redshifts(min_redshift: float, z_step_factor: float, max_redshift: float):
    """Log-spaced evolution nodes, highest first (wrapper/inputs.py:1774-1789)."""
    z = 10 ** np.arange(np.log10(1 + min_redshift), np.log10((1 + max_redshift) * z_step_factor),
                        np.log10(z_step_factor)) - 1
    return tuple(float(v) for v in z[::-1])


def run_coeval(*, out_redshifts=None, inputs: InputParameters, initial_conditions=None,
               backend: Backend | None = None):
    """ICs once, then perturb + ionize + T_b per redshift; returns a list of dicts, one per output
    redshift, highest first (drivers/coeval.py:530-700 without caching / halos / spin temperature).

    Without evolution every output redshift is independent.  With evolution (``RECOMB_MODEL`` set) the
    boxes are scrolled from the highest node down: the previous ionized box and perturbed field passed
    to each step are those of the last *node* redshift (``inputs.node_redshifts``, or log-spaced nodes
    from the lowest output redshift up to ``Z_HEAT_MAX`` in steps of ``ZPRIME_STEP_FACTOR``); output
    redshifts between nodes are computed from the node above them but do not feed the evolution
    (coeval.py:505-512)."""
    if out_redshifts is None:
        out_redshifts = inputs.node_redshifts
    outs = [float(out_redshifts)] if isinstance(out_redshifts, (int, float)) else [float(z) for z in out_redshifts]
    if not outs:
        raise ValueError("out_redshifts must be given if inputs has no node redshifts")
    if inputs.astro_options.USE_TS_FLUCT:
        raise NotImplementedError("run_coeval with USE_TS_FLUCT needs the spin-temperature calculation, which is "
                                  "outside the scoped path; pass a TsBox to compute_ionization_field instead")
    ics = initial_conditions or compute_initial_conditions(inputs=inputs, backend=backend)
    with _resident_ics(backend or get_backend()):
        return _run_coeval(outs, inputs, ics, backend)


class _resident_ics:
    """Keep the initial conditions in HBM for the duration of a driver loop that owns them (nothing
    mutates ``ics`` inside the loop): one upload instead of one per redshift -- and hand the perturbed density
    and the neutral fraction from one entry point to the next on the device (``b200_residency``: the loop
    below never modifies a box after the call that made it, and keeps the boxes of the current redshift alive
    while their copies are in use).  Both caches are opt-in in the library; other backends (the compiled
    reference) have no such symbols."""

    def __init__(self, backend):
        self.fns = []
        for name in ("b200_ics_cache", "b200_residency"):
            if hasattr(backend.lib, name):
                fn = getattr(backend.lib, name)
                fn.argtypes, fn.restype = [C.c_int], None
                self.fns.append(fn)

    def __enter__(self):
        for fn in self.fns:
            fn(1)

    def __exit__(self, *exc):
        for fn in self.fns:
            fn(0)
        return False


def _run_coeval(outs, inputs, ics, backend):
    out = []
    lagrangian = inputs.matter_options.lagrangian_source_grid

    def halobox(z):  # coeval.py:783-800: the source grids of this redshift (L-INTEGRAL: no halo catalogue)
        return compute_halobox(redshift=z, initial_conditions=ics, backend=backend) if lagrangian else None

    if not inputs.evolution_required:
        for z in sorted(outs, reverse=True):
            pf = perturb_field(redshift=z, initial_conditions=ics, backend=backend)
            ib = compute_ionization_field(perturbed_field=pf, initial_conditions=ics, halobox=halobox(z), backend=backend)
            bt = brightness_temperature(ionized_box=ib, perturbed_field=pf, backend=backend)
            out.append({"redshift": z, "perturbed_field": pf, "ionized_box": ib, "brightness_temp": bt})
        return out
    so = inputs.simulation_options
    nodes = tuple(inputs.node_redshifts) or get_logspaced_redshifts(min(outs), so.ZPRIME_STEP_FACTOR, so.Z_HEAT_MAX)
    prev_pf, prev_ib = PerturbedField.initial(inputs), IonizedBox.initial(inputs)
    for z in sorted(set(nodes) | set(outs), reverse=True):
        pf = perturb_field(redshift=z, initial_conditions=ics, backend=backend)
        ib = compute_ionization_field(perturbed_field=pf, initial_conditions=ics, previous_perturbed_field=prev_pf,
                                      previous_ionized_box=prev_ib, halobox=halobox(z), backend=backend)
        if z in outs:
            bt = brightness_temperature(ionized_box=ib, perturbed_field=pf, backend=backend)
            out.append({"redshift": z, "perturbed_field": pf, "ionized_box": ib, "brightness_temp": bt})
        if z in nodes:
            prev_pf, prev_ib = pf, ib
    return out


def run_coeval_parallel(*, out_redshifts, inputs: InputParameters, initial_conditions: InitialConditions,
                        backend: Backend | None = None, group=None):
    """The redshifts of a coeval run (or the snapshots of a lightcone) spread over the GPUs of a node, one process per
    GPU (``drivers/coeval.py:782-853`` loops over them on one device).  Every rank holds
    the same ``initial_conditions`` in host memory and takes the redshifts ``rank, rank + world, ...`` of the sorted
    list; the ranks call ``perturb_field`` in lockstep so that the upload of the initial conditions is shared
    (``SlabGroup.share_ics``: 1 / world of the arrays per PCIe link, the rest over NVLink).  Returns this rank's
    results, highest redshift first, as ``run_coeval`` does.  Needs options without evolution across redshifts."""
    import torch.distributed as dist

    from .distributed import SlabGroup
    if inputs.evolution_required:
        raise ValueError("redshifts are independent only without USE_TS_FLUCT, recombinations and mini-halos")
    be = backend or get_backend()
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    zs = sorted((float(z) for z in out_redshifts), reverse=True)
    rounds = (len(zs) + world - 1) // world
    lagrangian = inputs.matter_options.lagrangian_source_grid
    grp = SlabGroup(inputs=inputs, backend=be, group=group, heap_bytes=SlabGroup.ics_heap_bytes(inputs))
    residency = getattr(be.lib, "b200_residency", None) if hasattr(be.lib, "b200_residency") else None
    out = []
    try:
        grp.share_ics(world > 1)
        if residency is not None:  # density / neutral fraction stay on the device between the three calls of a redshift
            residency.argtypes, residency.restype = [C.c_int], None
            residency(1)
        for r in range(rounds):
            i = r * world + rank
            mine = i < len(zs)
            z = zs[i] if mine else zs[-1]  # a rank without work in the last round still takes part in the shared upload
            pf = perturb_field(redshift=z, initial_conditions=initial_conditions, backend=be)
            if not mine:
                continue
            hb = compute_halobox(redshift=z, initial_conditions=initial_conditions, backend=be) if lagrangian else None
            ib = compute_ionization_field(perturbed_field=pf, initial_conditions=initial_conditions, halobox=hb, backend=be)
            bt = brightness_temperature(ionized_box=ib, perturbed_field=pf, backend=be)
            out.append({"redshift": z, "perturbed_field": pf, "ionized_box": ib, "brightness_temp": bt})
    finally:
        if residency is not None:
            residency(0)
        grp.close()
    return out
