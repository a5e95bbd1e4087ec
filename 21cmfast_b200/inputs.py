"""Input parameter classes mirroring ``py21cmfast.wrapper.inputs`` for the grid hot path.

Same class names, field names, defaults and user-unit -> C-unit transformations as the reference
(``src/py21cmfast/wrapper/inputs.py``: CosmoParams :436-538, MatterOptions :642-801,
SimulationOptions :901-1084, AstroOptions :1184-1330, AstroParams :1427-1681, InputParameters
:1801+) so that parity tests read like the reference's own.  Only plain dataclasses are used
(the reference's attrs/astropy/classy machinery is out of scope); the C structs are built with the
ctypes layouts of ``_abi.py``.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from dataclasses import dataclass, field
from functools import cached_property

import numpy as np

from . import _abi

V_CB_AVG_DEFAULT = 27.0  # inputs.py:138


def _fill(struct_cls, values: dict):
    s = struct_cls()
    for name, _ in struct_cls._fields_:
        setattr(s, name, values[name])
    return s


class _InputStruct:
    """Shared helpers: ``clone``/``evolve``, ``asdict`` and a cached ``cstruct``."""

    def clone(self, **kw):
        return dataclasses.replace(self, **kw)

    evolve = clone

    def asdict(self):
        return dataclasses.asdict(self)

    @cached_property
    def cstruct(self):
        return _fill(self._struct_cls, self.cdict)


@dataclass(frozen=True)
class CosmoParams(_InputStruct):
    """Cosmological parameters (Planck18 defaults, inputs.py:492-538)."""

    SIGMA_8: float = 0.8102
    hlittle: float = 0.6766
    OMm: float = 0.30966
    OMb: float = 0.04897
    POWER_INDEX: float = 0.9665
    OMn: float = 0.0
    OMk: float = 0.0
    OMr: float = 8.6e-5
    OMtot: float = 1.0
    Y_He: float = 0.24
    wl: float = -1.0
    _struct_cls = _abi.CosmoParamsStruct

    @property
    def OMl(self):
        return 1 - self.OMm

    @property
    def cdict(self):
        d = {k: getattr(self, k) for k, _ in self._struct_cls._fields_ if k != "OMl"}
        d["OMl"] = self.OMl
        return d


@dataclass(frozen=True)
class SimulationOptions(_InputStruct):
    """Box geometry and stepping (inputs.py:1014-1084)."""

    HII_DIM: int = 256
    BOX_LEN: float | None = None
    DIM: int | None = None
    HIRES_TO_LOWRES_FACTOR: float | None = None
    LOWRES_CELL_SIZE_MPC: float | None = None
    NON_CUBIC_FACTOR: float = 1.0
    N_THREADS: int = 1
    SAMPLER_MIN_MASS: float = 1e8
    SAMPLER_BUFFER_FACTOR: float = 2.0
    N_COND_INTERP: int = 200
    N_PROB_INTERP: int = 400
    MIN_LOGPROB: float = -12
    HALOMASS_CORRECTION: float = 0.89
    PARKINSON_G0: float = 1.0
    PARKINSON_y1: float = 0.0
    PARKINSON_y2: float = 0.0
    Z_HEAT_MAX: float = 35.0
    ZPRIME_STEP_FACTOR: float = 1.02
    MIN_XE_FOR_FCOLL_IN_TAUX: float = 1e-3
    INITIAL_REDSHIFT: float = 300.0
    DELTA_R_FACTOR: float = 1.1
    DENSITY_SMOOTH_RADIUS: float = 0.2
    DEXM_OPTIMIZE_MINMASS: float = 1e11
    DEXM_R_OVERLAP: float = 2
    CORR_STAR: float = 0.5
    CORR_SFR: float = 0.2
    CORR_LX: float = 0.2
    _struct_cls = _abi.SimulationOptionsStruct

    def __post_init__(self):
        if self.DIM is not None and self.HIRES_TO_LOWRES_FACTOR is not None:
            raise ValueError("Cannot set both DIM and HIRES_TO_LOWRES_FACTOR!")
        if self.BOX_LEN is not None and self.LOWRES_CELL_SIZE_MPC is not None:
            raise ValueError("Cannot set both BOX_LEN and LOWRES_CELL_SIZE_MPC!")
        dcf, hdcf = self.dim * self.NON_CUBIC_FACTOR, self.HII_DIM * self.NON_CUBIC_FACTOR
        if dcf % int(dcf) or hdcf % int(hdcf):
            raise ValueError("NON_CUBIC_FACTOR * DIM and NON_CUBIC_FACTOR * HII_DIM must be integers")

    @property
    def dim(self) -> int:
        if self.DIM is not None:
            return int(self.DIM)
        f = 3 if self.HIRES_TO_LOWRES_FACTOR is None else self.HIRES_TO_LOWRES_FACTOR
        return int(self.HII_DIM * f)

    @property
    def box_len(self) -> float:
        if self.BOX_LEN is not None:
            return float(self.BOX_LEN)
        c = 1.5 if self.LOWRES_CELL_SIZE_MPC is None else self.LOWRES_CELL_SIZE_MPC
        return float(np.round(self.HII_DIM * c, 3))

    @property
    def HII_D_PARA(self) -> int:
        return int(self.NON_CUBIC_FACTOR * self.HII_DIM)

    @property
    def D_PARA(self) -> int:
        return int(self.NON_CUBIC_FACTOR * self.dim)

    @property
    def cdict(self):
        d = {k: getattr(self, k, None) for k, _ in self._struct_cls._fields_}
        d["DIM"] = self.dim
        d["BOX_LEN"] = self.box_len
        return d


@dataclass(frozen=True)
class MatterOptions(_InputStruct):
    """Matter-field options (inputs.py:766-801); strings map to the enum ints of InputParameters.h."""

    HMF: str = "ST"
    V_CB_MODEL: str = "NONE"
    POWER_SPECTRUM: str | None = None
    PERTURB_ON_HIGH_RES: bool = False
    USE_INTERPOLATION_TABLES: str = "hmf-interpolation"
    MINIMIZE_MEMORY: bool = False
    KEEP_3D_VELOCITIES: bool = False
    SAMPLE_METHOD: str = "MASS-LIMITED"
    FILTER: str = "spherical-tophat"
    HALO_FILTER: str = "spherical-tophat"
    SMOOTH_EVOLVED_DENSITY_FIELD: bool = False
    DEXM_OPTIMIZE: bool = False
    PERTURB_ALGORITHM: str = "2LPT"
    USE_FFTW_WISDOM: bool = False
    SOURCE_MODEL: str = "CHMF-SAMPLER"
    _struct_cls = _abi.MatterOptionsStruct

    def __post_init__(self):
        if self.FILTER == "sharp-k":
            raise ValueError("FILTER cannot be sharp-k")
        if self.V_CB_MODEL == "FLUCTS" and self.power_spectrum != "CLASS":
            raise ValueError("When using V_CB_MODEL='FLUCTS', you must use POWER_SPECTRUM = 'CLASS'!")

    @property
    def power_spectrum(self) -> str:
        if self.POWER_SPECTRUM is not None:
            return self.POWER_SPECTRUM
        return "CLASS" if self.V_CB_MODEL == "FLUCTS" else "EH"

    @property
    def lagrangian_source_grid(self) -> bool:
        return self.SOURCE_MODEL in ("L-INTEGRAL", "DEXM-ESF", "CHMF-SAMPLER")

    @property
    def mass_dependent_zeta(self) -> bool:
        return self.SOURCE_MODEL != "CONST-ION-EFF"

    @property
    def cdict(self):
        return dict(
            USE_FFTW_WISDOM=self.USE_FFTW_WISDOM, HMF=_abi.HMF[self.HMF],
            V_CB_MODEL=_abi.V_CB_MODEL[self.V_CB_MODEL],
            POWER_SPECTRUM=_abi.POWER_SPECTRUM[self.power_spectrum],
            USE_INTERPOLATION_TABLES=_abi.INTERPOLATION[self.USE_INTERPOLATION_TABLES],
            PERTURB_ON_HIGH_RES=self.PERTURB_ON_HIGH_RES,
            PERTURB_ALGORITHM=_abi.PERTURB_ALGORITHM[self.PERTURB_ALGORITHM],
            MINIMIZE_MEMORY=self.MINIMIZE_MEMORY, KEEP_3D_VELOCITIES=self.KEEP_3D_VELOCITIES,
            DEXM_OPTIMIZE=self.DEXM_OPTIMIZE, FILTER=_abi.FILTER[self.FILTER],
            HALO_FILTER=_abi.FILTER[self.HALO_FILTER],
            SMOOTH_EVOLVED_DENSITY_FIELD=self.SMOOTH_EVOLVED_DENSITY_FIELD,
            SOURCE_MODEL=_abi.SOURCE_MODEL[self.SOURCE_MODEL],
            SAMPLE_METHOD=_abi.SAMPLE_METHOD[self.SAMPLE_METHOD],
        )


@dataclass(frozen=True)
class AstroOptions(_InputStruct):
    """Astrophysical switches (inputs.py:1302-1330)."""

    USE_MINI_HALOS: bool = False
    USE_X_RAY_HEATING: bool = True
    USE_CMB_HEATING: bool = True
    USE_LYA_HEATING: bool = True
    RECOMB_MODEL: str = "none"
    USE_TS_FLUCT: bool = False
    USE_EXP_FILTER: bool = True
    CELL_RECOMB: bool = True
    LYA_MULTIPLE_SCATTERING: bool = False
    USE_ADIABATIC_FLUCTUATIONS: bool = True
    PHOTON_CONS_TYPE: str = "no-photoncons"
    USE_UPPER_STELLAR_TURNOVER: bool = True
    M_MIN_in_Mass: bool = True
    HALO_SCALING_RELATIONS_MEDIAN: bool = False
    HII_FILTER: str = "spherical-tophat"
    HEAT_FILTER: str = "spherical-tophat"
    IONISE_ENTIRE_SPHERE: bool = False
    INTEGRATION_METHOD_ATOMIC: str = "GAUSS-LEGENDRE"
    INTEGRATION_METHOD_MINI: str = "GAUSS-LEGENDRE"
    _struct_cls = _abi.AstroOptionsStruct

    def __post_init__(self):
        if self.USE_EXP_FILTER and self.HII_FILTER != "spherical-tophat":
            raise ValueError("USE_EXP_FILTER can only be used with a real-space tophat HII_FILTER==0")
        if self.USE_EXP_FILTER and not self.CELL_RECOMB:
            raise ValueError("USE_EXP_FILTER is True but CELL_RECOMB is False")
        if not self.CELL_RECOMB and self.RECOMB_MODEL == "homogeneous":  # wrapper/inputs.py:1384-1387
            raise ValueError("CELL_RECOMB cannot be False when RECOMB_MODEL is 'homogeneous'!")
        if self.USE_MINI_HALOS and (self.RECOMB_MODEL == "none" or not self.USE_TS_FLUCT):
            raise ValueError("USE_MINI_HALOS needs RECOMB_MODEL != 'none' and USE_TS_FLUCT")

    @property
    def cdict(self):
        d = {k: getattr(self, k) for k, _ in self._struct_cls._fields_}
        d["RECOMB_MODEL"] = _abi.RECOMB_MODEL[self.RECOMB_MODEL]
        d["PHOTON_CONS_TYPE"] = _abi.PHOTON_CONS[self.PHOTON_CONS_TYPE]
        d["HII_FILTER"] = _abi.FILTER[self.HII_FILTER]
        d["HEAT_FILTER"] = _abi.FILTER[self.HEAT_FILTER]
        d["INTEGRATION_METHOD_ATOMIC"] = _abi.INTEGRATION_METHOD[self.INTEGRATION_METHOD_ATOMIC]
        d["INTEGRATION_METHOD_MINI"] = _abi.INTEGRATION_METHOD[self.INTEGRATION_METHOD_MINI]
        return d


_LOG10_FIELDS = ("F_STAR10", "F_STAR7_MINI", "F_ESC10", "F_ESC7_MINI", "M_TURN", "ION_Tvir_MIN",
                 "L_X", "L_X_MINI", "X_RAY_Tvir_MIN", "UPPER_STELLAR_TURNOVER_MASS")
_DEX_FIELDS = ("SIGMA_STAR", "SIGMA_LX", "SIGMA_SFR_LIM", "SIGMA_SFR_INDEX")


@dataclass(frozen=True)
class AstroParams(_InputStruct):
    """Astrophysical parameters in user units (log10 where the reference uses log10);
    ``cdict`` applies the reference's transformers (inputs.py:1569-1681)."""

    HII_EFF_FACTOR: float = 30.0
    F_STAR10: float = -1.3
    ALPHA_STAR: float = 0.5
    F_STAR7_MINI: float | None = None
    ALPHA_STAR_MINI: float | None = None
    F_ESC10: float = -1.0
    ALPHA_ESC: float = -0.5
    F_ESC7_MINI: float = -2.0
    M_TURN: float = 8.7
    R_BUBBLE_MAX: float = 15.0
    R_BUBBLE_MIN: float = 0.620350491
    ION_Tvir_MIN: float = 4.69897
    L_X: float = 40.5
    L_X_MINI: float | None = None
    NU_X_THRESH: float = 500.0
    X_RAY_SPEC_INDEX: float = 1.0
    X_RAY_Tvir_MIN: float | None = None
    F_H2_SHIELD: float = 0.0
    t_STAR: float = 0.5
    A_LW: float = 2.0
    BETA_LW: float = 0.6
    A_VCB: float = 1.0
    BETA_VCB: float = 1.8
    UPPER_STELLAR_TURNOVER_MASS: float = 11.447
    UPPER_STELLAR_TURNOVER_INDEX: float = -0.6
    SIGMA_STAR: float = 0.25
    SIGMA_LX: float = 0.5
    SIGMA_SFR_LIM: float = 0.19
    SIGMA_SFR_INDEX: float = -0.12
    T_RE: float = 2e4
    V_CB_AVG_DEBUG: float = V_CB_AVG_DEFAULT
    POP2_ION: float = 5000.0
    POP3_ION: float = 44021.0
    PHOTONCONS_CALIBRATION_END: float = 3.5
    CLUMPING_FACTOR: float = 2.0
    ALPHA_UVB: float = 5.0
    R_MAX_TS: float = 500.0
    N_STEP_TS: int = 40
    MAX_DVDR: float = 0.2
    DELTA_R_HII_FACTOR: float = 1.1
    NU_X_BAND_MAX: float = 2000.0
    NU_X_MAX: float = 10000.0
    _struct_cls = _abi.AstroParamsStruct

    def _user(self, name):
        v = getattr(self, name)
        if v is not None:
            return v
        return {"F_STAR7_MINI": self.F_STAR10 - 3 * self.ALPHA_STAR,
                "ALPHA_STAR_MINI": self.ALPHA_STAR, "L_X_MINI": self.L_X,
                "X_RAY_Tvir_MIN": self.ION_Tvir_MIN}[name]

    @property
    def cdict(self):
        d = {}
        for k, _ in self._struct_cls._fields_:
            v = self._user(k)
            if k in _LOG10_FIELDS:
                v = 10.0 ** v
            elif k in _DEX_FIELDS:
                v = v * math.log(10.0)
            d[k] = v
        return d


@dataclass(frozen=True, eq=False)
class CosmoTables:
    """Derived tables (inputs.py:358-382).  With ``POWER_SPECTRUM='CLASS'`` the reference fills
    ``transfer_density`` / ``transfer_vcb`` by running ``classy`` (inputs.py:1864-1966); that package is not
    part of this build, so the caller passes the ``(k [1/Mpc], T(k))`` samples (k = 0 first, as the reference
    prepends it) -- from a classy run elsewhere, a file, or any tabulated transfer function."""

    ps_norm: float
    USE_SIGMA_8: bool = True
    V_CB_AVG: float = V_CB_AVG_DEFAULT
    transfer_density: tuple | None = None  # (k, T_m(k, z=0))
    transfer_vcb: tuple | None = None      # (k, T_vcb(k, z_dec) / c), same k

    def _table(self, kt):
        k = np.ascontiguousarray(kt[0], dtype=np.float64)
        t = np.ascontiguousarray(kt[1], dtype=np.float64)
        if k.ndim != 1 or k.shape != t.shape or k.size < 3:
            raise ValueError("a transfer table is a pair of equally long 1-D arrays (k, T)")
        tab = _abi.Table1DStruct()
        tab.size = k.size
        tab.x_values = k.ctypes.data_as(_abi.c_double_p)
        tab.y_values = t.ctypes.data_as(_abi.c_double_p)
        return tab, (k, t)

    @cached_property
    def cstruct(self):
        s = _abi.CosmoTablesStruct()
        keep = []
        for name in ("transfer_density", "transfer_vcb"):
            kt = getattr(self, name)
            if kt is None:
                setattr(s, name, None)
            else:
                tab, arrays = self._table(kt)
                keep.append((tab, arrays))
                setattr(s, name, C.pointer(tab))
        s.ps_norm = self.ps_norm
        s.USE_SIGMA_8 = self.USE_SIGMA_8
        s.V_CB_AVG = self.V_CB_AVG
        s._keep = keep  # the C side deep-copies the tables at broadcast (InputParameters.c:21-53)
        return s


@dataclass(frozen=True)
class InputParameters:
    """Bundle of all input structs plus the seed (inputs.py:1801-1840)."""

    random_seed: int
    cosmo_params: CosmoParams = field(default_factory=CosmoParams)
    simulation_options: SimulationOptions = field(default_factory=SimulationOptions)
    matter_options: MatterOptions = field(default_factory=MatterOptions)
    astro_params: AstroParams = field(default_factory=AstroParams)
    astro_options: AstroOptions = field(default_factory=AstroOptions)
    node_redshifts: tuple = ()
    class_tables: CosmoTables | None = None  # POWER_SPECTRUM='CLASS': the tables classy would have made

    def __post_init__(self):
        if self.matter_options.power_spectrum == "CLASS":
            ct = self.class_tables
            if ct is None or ct.transfer_density is None:
                raise NotImplementedError("POWER_SPECTRUM='CLASS' needs classy, which is not part of this build: pass "
                                          "class_tables=CosmoTables(transfer_density=(k, T), ...)")
            if self.matter_options.V_CB_MODEL == "FLUCTS" and ct.transfer_vcb is None:
                raise ValueError("V_CB_MODEL='FLUCTS' needs class_tables.transfer_vcb")

    @cached_property
    def cosmo_tables(self) -> CosmoTables:
        if self.matter_options.power_spectrum == "CLASS":
            return self.class_tables
        return CosmoTables(ps_norm=self.cosmo_params.SIGMA_8, USE_SIGMA_8=True)

    @property
    def evolution_required(self) -> bool:
        return (self.astro_options.USE_TS_FLUCT or self.astro_options.RECOMB_MODEL != "none"
                or self.astro_options.USE_MINI_HALOS)

    def evolve_input_structs(self, **kwargs):
        """Return a copy with individual fields of any sub-struct replaced (inputs.py API)."""
        subs = {}
        for name in ("cosmo_params", "simulation_options", "matter_options", "astro_params",
                     "astro_options"):
            s = getattr(self, name)
            names = {f.name for f in dataclasses.fields(s)}
            upd = {k: v for k, v in kwargs.items() if k in names}
            subs[name] = s.clone(**upd) if upd else s
        used = set()
        for name in subs:
            used |= {f.name for f in dataclasses.fields(getattr(self, name))} & set(kwargs)
        extra = set(kwargs) - used
        if extra:
            raise TypeError(f"unknown parameter(s): {sorted(extra)}")
        return dataclasses.replace(self, **subs)

    def clone(self, **kw):
        return dataclasses.replace(self, **kw)
