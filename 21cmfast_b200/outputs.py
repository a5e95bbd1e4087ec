"""Output structs mirroring ``py21cmfast.wrapper.outputs`` for the grid hot path.

numpy owns every array (``outputs.py:4-7`` in the reference); the C side receives raw pointers
and fills them.  Shapes, dtypes, optional arrays and initial values follow
``InitialConditions.new`` (outputs.py:533-582), ``PerturbedField.new`` (:688-719) and
``IonizedBox.new`` (:1475-1545; ``neutral_fraction`` starts at one).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .inputs import InputParameters


def _ptr(a):
    return a.ctypes.data_as(_abi.c_float_p) if a is not None else None


class OutputStruct:
    _struct_cls = None
    _arrays = ()
    _scalars = ()

    def __init__(self, inputs: InputParameters, **arrays):
        self.inputs = inputs
        self.is_computed = False
        for k in self._arrays:
            setattr(self, k, arrays.get(k))
        for k in self._scalars:
            setattr(self, k, 0.0)
        self._c = None

    @property
    def cstruct(self):
        """Fresh C struct with the current numpy data pointers (structs.py:80-95)."""
        s = self._struct_cls()
        for k in self._arrays:
            a = getattr(self, k)
            if a is not None:
                assert a.dtype == np.float32 and a.flags.c_contiguous
            setattr(s, k, _ptr(a))
        for k in self._scalars:
            setattr(s, k, getattr(self, k))
        self._c = s
        return s

    def pull_scalars(self):
        for k in self._scalars:
            setattr(self, k, getattr(self._c, k))

    def arrays(self):
        return {k: getattr(self, k) for k in self._arrays if getattr(self, k) is not None}


def _shapes(inputs):
    so = inputs.simulation_options
    return (so.HII_DIM, so.HII_DIM, so.HII_D_PARA), (so.dim, so.dim, so.D_PARA)


class InitialConditions(OutputStruct):
    _struct_cls = _abi.InitialConditionsStruct
    _arrays = tuple(n for n, _ in _abi.InitialConditionsStruct._fields_)

    @classmethod
    def new(cls, inputs: InputParameters):
        lo, hi = _shapes(inputs)
        mo = inputs.matter_options
        z = lambda s: np.zeros(s, dtype=np.float32)  # noqa: E731
        out = {"lowres_density": z(lo), "hires_density": z(hi)}
        if mo.PERTURB_ON_HIGH_RES:
            out |= {f"hires_v{a}": z(hi) for a in "xyz"}
        else:
            out |= {f"lowres_v{a}": z(lo) for a in "xyz"}
        if mo.PERTURB_ALGORITHM == "2LPT":
            out |= {f"hires_v{a}_2LPT": z(hi) for a in "xyz"}
            if not mo.PERTURB_ON_HIGH_RES:
                out |= {f"lowres_v{a}_2LPT": z(lo) for a in "xyz"}
        if mo.V_CB_MODEL == "FLUCTS":
            out["lowres_vcb"] = z(lo)
        return cls(inputs, **out)


class PerturbedField(OutputStruct):
    _struct_cls = _abi.PerturbedFieldStruct
    _arrays = ("density", "velocity_x", "velocity_y", "velocity_z")

    def __init__(self, inputs, redshift, **kw):
        super().__init__(inputs, **kw)
        self.redshift = redshift

    @classmethod
    def new(cls, inputs: InputParameters, redshift: float):
        lo, _ = _shapes(inputs)
        out = {"density": np.zeros(lo, np.float32), "velocity_z": np.zeros(lo, np.float32)}
        if inputs.matter_options.KEEP_3D_VELOCITIES:
            out["velocity_x"] = np.zeros(lo, np.float32)
            out["velocity_y"] = np.zeros(lo, np.float32)
        return cls(inputs, redshift, **out)

    @classmethod
    def initial(cls, inputs):
        """The 'initial' previous box: zeros, redshift -1 (single_field.py:778-791)."""
        return cls.new(inputs, redshift=-1.0)


class TsBox(OutputStruct):
    _struct_cls = _abi.TsBoxStruct
    _arrays = ("spin_temperature", "xray_ionised_fraction", "kinetic_temp_neutral", "J_21_LW")
    _scalars = ("Q_HI",)

    def __init__(self, inputs, redshift=None, **kw):
        super().__init__(inputs, **kw)
        self.redshift = redshift

    @classmethod
    def dummy(cls, inputs):
        return cls(inputs)

    @classmethod
    def new(cls, inputs: InputParameters, redshift: float):
        """Allocated arrays of a spin-temperature box (outputs.py ``TsBox.new``).  The spin-temperature
        calculation itself is outside the scoped path: callers fill the arrays (from the reference, or
        from a file) and pass the box to ``compute_ionization_field`` / ``brightness_temperature``."""
        lo, _ = _shapes(inputs)
        out = {k: np.zeros(lo, np.float32) for k in ("spin_temperature", "xray_ionised_fraction",
                                                     "kinetic_temp_neutral")}
        if inputs.astro_options.USE_MINI_HALOS:
            out["J_21_LW"] = np.zeros(lo, np.float32)
        return cls(inputs, redshift, **out)


class HaloBox(OutputStruct):
    _struct_cls = _abi.HaloBoxStruct
    _arrays = tuple(n for n, t in _abi.HaloBoxStruct._fields_ if t is _abi.c_float_p)
    _scalars = ("log10_Mcrit_ACG_ave", "log10_Mcrit_MCG_ave")

    def __init__(self, inputs, redshift=None, **kw):
        super().__init__(inputs, **kw)
        self.redshift = redshift

    @classmethod
    def dummy(cls, inputs):
        return cls(inputs)

    @classmethod
    def new(cls, inputs: InputParameters, redshift: float):
        """Arrays of ``HaloBox.new`` (outputs.py:1095-1135) for the options in scope: the photon-output and
        star-formation grids, plus the escape-weighted star formation with recombinations."""
        lo, _ = _shapes(inputs)
        out = {"halo_sfr": np.zeros(lo, np.float32), "n_ion": np.zeros(lo, np.float32)}
        if inputs.astro_options.RECOMB_MODEL != "none":
            out["whalo_sfr"] = np.zeros(lo, np.float32)
        return cls(inputs, redshift, **out)


class IonizedBox(OutputStruct):
    _struct_cls = _abi.IonizedBoxStruct
    _arrays = ("neutral_fraction", "ionisation_rate_G12", "mean_free_path", "z_reion",
               "cumulative_recombinations", "kinetic_temperature", "unnormalised_nion",
               "unnormalised_nion_mini")
    _scalars = ("mean_f_coll", "mean_f_coll_MINI", "log10_Mturnover_ave",
                "log10_Mturnover_MINI_ave")

    def __init__(self, inputs, redshift, **kw):
        super().__init__(inputs, **kw)
        self.redshift = redshift

    @classmethod
    def new(cls, inputs: InputParameters, redshift: float):
        lo, _ = _shapes(inputs)
        ap, ao, mo, so = (inputs.astro_params, inputs.astro_options, inputs.matter_options,
                          inputs.simulation_options)
        n_filtering = 1
        if ao.USE_MINI_HALOS and not mo.lagrangian_source_grid and so.HII_DIM > 1:
            n_filtering = int(np.log(min(ap.R_BUBBLE_MAX, 0.620350491 * so.box_len)
                                     / max(ap.R_BUBBLE_MIN, 0.620350491 * so.box_len / so.HII_DIM))
                              / np.log(ap.DELTA_R_HII_FACTOR)) + 1
        z = lambda s: np.zeros(s, dtype=np.float32)  # noqa: E731
        out = {"neutral_fraction": np.ones(lo, np.float32), "ionisation_rate_G12": z(lo),
               "z_reion": z(lo)}
        if not mo.MINIMIZE_MEMORY:
            out["mean_free_path"] = z(lo)
            out["kinetic_temperature"] = z(lo)
        if ao.RECOMB_MODEL == "inhomogeneous":
            out["cumulative_recombinations"] = z(lo)
        elif ao.RECOMB_MODEL == "homogeneous":
            out["cumulative_recombinations"] = z((1, 1, 1))
        if not mo.lagrangian_source_grid:
            out["unnormalised_nion"] = z((n_filtering, *lo))
            if ao.USE_MINI_HALOS:
                out["unnormalised_nion_mini"] = z((n_filtering, *lo))
        return cls(inputs, redshift, **out)

    @classmethod
    def initial(cls, inputs):
        return cls.new(inputs, redshift=-1.0)

    @property
    def global_xH(self):
        return float(np.mean(self.neutral_fraction, dtype=np.float64))


class BrightnessTemp(OutputStruct):
    _struct_cls = _abi.BrightnessTempStruct
    _arrays = ("brightness_temp", "tau_21")

    def __init__(self, inputs, redshift, **kw):
        super().__init__(inputs, **kw)
        self.redshift = redshift

    @classmethod
    def new(cls, inputs, redshift):
        lo, _ = _shapes(inputs)
        return cls(inputs, redshift, brightness_temp=np.zeros(lo, np.float32),
                   tau_21=np.zeros(lo, np.float32))
