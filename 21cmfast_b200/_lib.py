"""Loader for the C-ABI shared library and the global-state manager.

``Backend`` binds any shared object that exports the reference's C symbols
(``src/py21cmfast/src/_functionprototypes_wrapper.h``): by default the B200 library
``lib21cmfast_b200.so`` built from ``csrc/``.  Loading fails loudly if the library is missing --
there is no CPU fallback on the product path.  (Tests bind the compiled reference in
``oracle/_ref`` through the same class to use it as the checker.)

``GlobalState`` mirrors ``GlobalInitializationManager``
(``src/py21cmfast/drivers/_global_initialization.py:18-159``): parameters are passed by pointer to
Python-owned structs via ``Broadcast_struct_global_all`` and the per-input tables are
initialised lazily (``init_ps``, ``initialiseSigmaMInterpTable(5e2, 1e20)``, ``init_heat``).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import _abi

_HERE = Path(__file__).resolve().parent
DEFAULT_LIB = _HERE / "csrc" / "lib21cmfast_b200.so"


class BackendError(RuntimeError):
    """Raised for a non-zero status from a ``Compute*`` call (exceptions.h:12-21)."""

    def __init__(self, code, where):
        self.code = code
        super().__init__(f"{where} failed with status {code} ({_abi.ERROR_CODES.get(code, 'Unknown error in C')})")


class Backend:
    def __init__(self, path: str | os.PathLike | None = None):
        path = Path(path) if path is not None else DEFAULT_LIB
        if not path.exists():
            raise ImportError(
                f"21cmfast_b200: C-ABI library {path} not found. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU fallback.")
        self.path = path
        self.lib = lib = C.CDLL(str(path), mode=C.RTLD_LOCAL)
        P = C.POINTER
        lib.ComputeInitialConditions.argtypes = [C.c_ulonglong, P(_abi.InitialConditionsStruct)]
        lib.ComputePerturbedField.argtypes = [C.c_float, P(_abi.InitialConditionsStruct),
                                              P(_abi.PerturbedFieldStruct)]
        lib.ComputeIonizedBox.argtypes = [
            C.c_float, C.c_float, P(_abi.PerturbedFieldStruct), P(_abi.PerturbedFieldStruct),
            P(_abi.IonizedBoxStruct), P(_abi.TsBoxStruct), P(_abi.HaloBoxStruct),
            P(_abi.InitialConditionsStruct), P(_abi.IonizedBoxStruct)]
        lib.ComputeBrightnessTemp.argtypes = [
            C.c_float, P(_abi.TsBoxStruct), P(_abi.IonizedBoxStruct), P(_abi.PerturbedFieldStruct),
            P(_abi.BrightnessTempStruct)]
        lib.test_filter.argtypes = [C.POINTER(C.c_float), C.c_double, C.c_double, C.c_double,
                                    C.c_int, C.POINTER(C.c_double)]
        for name in ("ComputeInitialConditions", "ComputePerturbedField", "ComputeIonizedBox",
                     "ComputeBrightnessTemp", "test_filter", "init_heat", "CreateFFTWWisdoms"):
            getattr(lib, name).restype = C.c_int
        lib.Broadcast_struct_global_all.argtypes = [
            P(_abi.SimulationOptionsStruct), P(_abi.MatterOptionsStruct),
            P(_abi.CosmoParamsStruct), P(_abi.AstroParamsStruct), P(_abi.AstroOptionsStruct),
            P(_abi.CosmoTablesStruct)]
        lib.Broadcast_struct_global_all.restype = None
        lib.initialiseSigmaMInterpTable.argtypes = [C.c_float, C.c_float]
        for name in ("init_ps", "free_ps", "initialiseSigmaMInterpTable", "freeSigmaMInterpTable",
                     "destruct_heat", "Free_cosmo_tables_global", "init_MHR", "free_MHR"):
            getattr(lib, name).restype = None
        lib.splined_recombination_rate.argtypes = [C.c_double, C.c_double]
        lib.splined_recombination_rate.restype = C.c_double
        for name, args in (("dicke", [C.c_double]), ("sigma_z0", [C.c_double]),
                           ("dsigmasqdm_z0", [C.c_double]), ("power_in_k", [C.c_double])):
            f = getattr(lib, name)
            f.argtypes, f.restype = args, C.c_double
        self.config = _abi.ConfigSettingsStruct.in_dll(lib, "config_settings")
        self.state = GlobalState(self)
        self._table_path = None

    def ensure_table_path(self):
        """Point ``config_settings.external_table_path`` at the packaged RECFAST table unless the
        caller already chose a directory with ``set_table_path`` (the reference sets it from
        ``py21cmfast/_cfg.py:61-70`` at import time)."""
        if self._table_path is None:
            from ._data import default_table_dir
            self.set_table_path(default_table_dir())

    def set_table_path(self, path):
        self._table_path = os.fsencode(str(path))
        self.config.external_table_path = self._table_path

    def set_wisdoms_path(self, path):
        self._wisdoms_path = os.fsencode(str(path))
        self.config.wisdoms_path = self._wisdoms_path


class GlobalState:
    def __init__(self, backend: Backend):
        self.backend = backend
        self.inputs = None
        self._keep = None
        self.inputs_are_broadcast = self.ps_inited = self.sigma_inited = self.heat_inited = False
        self.recomb_inited = False

    def free(self):
        lib = self.backend.lib
        if self.recomb_inited:
            lib.free_MHR()
            self.recomb_inited = False
        if self.heat_inited:
            lib.destruct_heat()
            self.heat_inited = False
        if self.sigma_inited:
            lib.freeSigmaMInterpTable()
            self.sigma_inited = False
        if self.ps_inited:
            lib.free_ps()
            self.ps_inited = False
        if self.inputs_are_broadcast:
            lib.Free_cosmo_tables_global()
            self.inputs_are_broadcast = False

    def init(self, inputs, *, broadcast_inputs=False, ps=False, sigma=False, heat=False, recomb=False):
        lib = self.backend.lib
        if self.inputs is None or self.inputs != inputs:
            self.free()
            self.inputs = inputs
        i = self.inputs
        if (broadcast_inputs or ps or sigma or heat or recomb) and not self.inputs_are_broadcast:
            if hasattr(lib, "b200_set_device"):
                # another binding of the same shared object (cffi, a second Backend) may have left its copy of the
                # cosmo tables behind: the C side copies them only once per Free_cosmo_tables_global
                # (InputParameters.c:9-53); the library's own free is a no-op when nothing is allocated
                lib.Free_cosmo_tables_global()
            # keep the structs alive: the C side stores *pointers* (InputParameters.c:11-20)
            self._keep = (i.simulation_options.cstruct, i.matter_options.cstruct,
                          i.cosmo_params.cstruct, i.astro_params.cstruct,
                          i.astro_options.cstruct, i.cosmo_tables.cstruct)
            lib.Broadcast_struct_global_all(*[C.byref(s) for s in self._keep])
            if i.matter_options.USE_FFTW_WISDOM:
                lib.CreateFFTWWisdoms()
            self.inputs_are_broadcast = True
        if (ps or sigma) and not self.ps_inited:
            lib.init_ps()
            self.ps_inited = True
        if (sigma and not self.sigma_inited
                and self.inputs.matter_options.USE_INTERPOLATION_TABLES != "no-interpolation"):
            lib.initialiseSigmaMInterpTable(5e2, 1e20)
            self.sigma_inited = True
        if (recomb and not self.recomb_inited and i.astro_options.RECOMB_MODEL != "none"
                and i.simulation_options.HII_DIM > 1):
            lib.init_MHR()  # _global_initialization.py:147-159
            self.recomb_inited = True
        if heat and not self.heat_inited:
            self.backend.ensure_table_path()
            status = lib.init_heat()
            if status != 0:
                raise BackendError(status, "init_heat")
            self.heat_inited = True


_default_backend = None


def get_backend() -> Backend:
    """The process-wide B200 backend (loads the CUDA library on first use)."""
    global _default_backend
    if _default_backend is None:
        _default_backend = Backend()
    return _default_backend
