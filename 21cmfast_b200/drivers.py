"""Single-field drivers mirroring ``py21cmfast.drivers.single_field`` for the hot path.

``compute_initial_conditions`` (single_field.py:38-113), ``perturb_field`` (:116-156),
``compute_ionization_field`` (:714-840), ``brightness_temperature`` and a minimal ``run_coeval``
(coeval.py:521-697, evolution-free configs only).  Each call initialises the backend's global
state exactly as ``@init_c_state`` does in the reference and then calls the C-ABI entry point
with numpy-owned host buffers.
"""
from __future__ import annotations

import ctypes as C

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


def compute_ionization_field(*, perturbed_field: PerturbedField,
                             initial_conditions: InitialConditions,
                             previous_perturbed_field: PerturbedField | None = None,
                             previous_ionized_box: IonizedBox | None = None,
                             spin_temp: TsBox | None = None,
                             backend: Backend | None = None) -> IonizedBox:
    be = backend or get_backend()
    inputs = perturbed_field.inputs
    ao = inputs.astro_options
    if ao.USE_MINI_HALOS or inputs.matter_options.lagrangian_source_grid:
        raise NotImplementedError(
            "only the Eulerian IonizeBox path without mini-halos is in scope (SURVEY.md section 8)")
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
    ts, hb = (spin_temp if ao.USE_TS_FLUCT else TsBox.dummy(inputs)), HaloBox.dummy(inputs)
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


def run_coeval(*, out_redshifts, inputs: InputParameters, initial_conditions=None,
               backend: Backend | None = None):
    """ICs once, then perturb + ionize (+ T_b) per redshift; returns a list of dicts."""
    ics = initial_conditions or compute_initial_conditions(inputs=inputs, backend=backend)
    out = []
    for z in ([out_redshifts] if isinstance(out_redshifts, (int, float)) else out_redshifts):
        pf = perturb_field(redshift=z, initial_conditions=ics, backend=backend)
        ib = compute_ionization_field(perturbed_field=pf, initial_conditions=ics, backend=backend)
        out.append({"redshift": z, "perturbed_field": pf, "ionized_box": ib})
    return out
