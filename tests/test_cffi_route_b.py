"""INTEGRATION.md route B for real: ``cffi`` ABI mode -- ``ffi.cdef`` of the reference's own wrapper
headers (where /root/reference exists; else of include/py21cmfast_b200.h, which test_abi.py proves
layout-identical) + ``ffi.dlopen`` of the shipped library -- and calls through the resulting
``lib``, exactly as ``py21cmfast.c_21cmfast`` would be replaced.  The CPU tier exercises the host
entry points (parameter broadcast, power spectrum, sigma(M), growth); the GPU tier runs
ComputePerturbedField / ComputeIonizedBox through cffi and compares with the golden fixtures."""
import re
from pathlib import Path

import numpy as np
import pytest

import common

pkg = common.pkg
ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "21cmfast_b200" / "csrc" / "lib21cmfast_b200.so"
REF_SRC = Path("/root/reference/src/py21cmfast/src")


def _cdef_text():
    if REF_SRC.exists():  # the text build_cffi.py:166-178 feeds to ffi.cdef
        return [(REF_SRC / h).read_text() for h in
                ("_inputparams_wrapper.h", "_outputstructs_wrapper.h", "_functionprototypes_wrapper.h")], "reference"
    txt = (ROOT / "include" / "py21cmfast_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    lines = [ln for ln in txt.splitlines() if not ln.lstrip().startswith("#") and 'extern "C"' not in ln]
    body = "\n".join(lines)
    body = body[:body.rindex("}")] if body.rstrip().endswith("}") else body  # closing brace of extern "C"
    return [body], "include/py21cmfast_b200.h"


@pytest.fixture(scope="module")
def cffi_lib():
    cffi = pytest.importorskip("cffi")
    if not LIB.exists():
        pytest.skip("product library not built")
    ffi = cffi.FFI()
    texts, origin = _cdef_text()
    for t in texts:
        ffi.cdef(t)
    ffi.cdef("void free(void *ptr);")
    lib = ffi.dlopen(str(LIB))
    return ffi, lib, origin


def _broadcast(ffi, lib, inputs):
    """what GlobalInitializationManager does (drivers/_global_initialization.py:63-100) with cffi structs"""
    keep = []
    ptrs = []
    for cname, sub in (("SimulationOptions", inputs.simulation_options), ("MatterOptions", inputs.matter_options),
                       ("CosmoParams", inputs.cosmo_params), ("AstroParams", inputs.astro_params),
                       ("AstroOptions", inputs.astro_options)):
        s = ffi.new(f"{cname} *")
        for k, v in sub.cdict.items():
            setattr(s, k, v)
        keep.append(s)
        ptrs.append(s)
    ct = ffi.new("CosmoTables *")
    ct.transfer_density = ffi.NULL
    ct.transfer_vcb = ffi.NULL
    ct.ps_norm = inputs.cosmo_tables.ps_norm
    ct.USE_SIGMA_8 = inputs.cosmo_tables.USE_SIGMA_8
    ct.V_CB_AVG = inputs.cosmo_tables.V_CB_AVG
    keep.append(ct)
    lib.Broadcast_struct_global_all(*ptrs, ct)
    return keep


def test_cffi_dlopen_host_entry_points(cffi_lib):
    ffi, lib, origin = cffi_lib
    inputs = common.make_inputs(hii=32, dim=64)
    keep = _broadcast(ffi, lib, inputs)
    lib.init_ps()
    lib.initialiseSigmaMInterpTable(5e2, 1e20)
    # the same scalars through the ctypes binding the rest of the suite uses
    be = pkg.Backend(LIB)
    be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True)
    for z in (6.0, 8.0, 25.0):
        assert lib.dicke(z) == be.lib.dicke(z)
    for m in (1e8, 1e10, 1e13):
        assert lib.sigma_z0(m) == be.lib.sigma_z0(m)
    assert 0.1 < lib.dicke(8.0) < 0.2
    g = np.load(common.GOLDEN / "host_scalars.npz")
    if "dicke_z8" in g:
        assert abs(lib.dicke(8.0) - float(g["dicke_z8"])) < 1e-14
    # the globals are visible through cffi too
    assert lib.simulation_options_global.HII_DIM == 32
    assert lib.matter_options_global.SOURCE_MODEL == inputs.matter_options.cdict["SOURCE_MODEL"]
    lib.freeSigmaMInterpTable()
    lib.free_ps()
    lib.Free_cosmo_tables_global()  # _global_initialization.py:77-87 frees the C-side copy with the structs
    del keep
    print("cdef from", origin)


@pytest.mark.gpu
def test_cffi_dlopen_compute_calls_reproduce_golden(cffi_lib):
    ffi, lib, origin = cffi_lib
    inputs, ics, g_pf, g_ib = common.load_golden(common.GOLDEN_BASE)
    keep = _broadcast(ffi, lib, inputs)
    path = ffi.new("char[]", str(common.table_dir()).encode())
    lib.config_settings.external_table_path = path
    lib.init_ps()
    lib.initialiseSigmaMInterpTable(5e2, 1e20)
    assert lib.init_heat() == 0

    def fptr(a):
        return ffi.cast("float *", ffi.from_buffer(a)) if a is not None else ffi.NULL

    c_ics = ffi.new("InitialConditions *")
    for k in ics._arrays:
        setattr(c_ics, k, fptr(getattr(ics, k)))
    pf = pkg.PerturbedField.new(inputs, 8.0)
    c_pf = ffi.new("PerturbedField *")
    for k in pf._arrays:
        setattr(c_pf, k, fptr(getattr(pf, k)))
    assert lib.ComputePerturbedField(8.0, c_ics, c_pf) == 0
    common.compare_struct(pf, g_pf, tols={"velocity_z": common.TOL_VELOCITY})

    ib = pkg.IonizedBox.new(inputs, 8.0)
    prev_ib, prev_pf = pkg.IonizedBox.initial(inputs), pkg.PerturbedField.initial(inputs)
    c_ib, c_prev_ib = ffi.new("IonizedBox *"), ffi.new("IonizedBox *")
    c_prev_pf, c_ts, c_hb = ffi.new("PerturbedField *"), ffi.new("TsBox *"), ffi.new("HaloBox *")
    for obj, c in ((ib, c_ib), (prev_ib, c_prev_ib), (prev_pf, c_prev_pf)):
        for k in obj._arrays:
            setattr(c, k, fptr(getattr(obj, k)))
    c_gpf = ffi.new("PerturbedField *")
    for k in g_pf._arrays:
        setattr(c_gpf, k, fptr(getattr(g_pf, k)))
    assert lib.ComputeIonizedBox(8.0, -1.0, c_gpf, c_prev_pf, c_prev_ib, c_ts, c_hb, c_ics, c_ib) == 0
    ib.mean_f_coll, ib.log10_Mturnover_ave = c_ib.mean_f_coll, c_ib.log10_Mturnover_ave
    stats = common.compare_ionized(ib, g_ib)
    assert stats["mask_mismatch"] == 0
    lib.destruct_heat()
    lib.freeSigmaMInterpTable()
    lib.free_ps()
    lib.Free_cosmo_tables_global()
    del keep
