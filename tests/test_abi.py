"""The C-ABI contract: struct layouts of include/py21cmfast_b200.h equal the ctypes mirror (and,
where /root/reference is present, the reference's own headers), and the built library exports
every function the header declares.  CPU only (no compute calls)."""
import ctypes as C
import re
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

import common

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "py21cmfast_b200.h"
LIB = ROOT / "21cmfast_b200" / "csrc" / "lib21cmfast_b200.so"
_abi = __import__("importlib").import_module("21cmfast_b200._abi")

STRUCTS = {
    "CosmoParams": _abi.CosmoParamsStruct, "SimulationOptions": _abi.SimulationOptionsStruct,
    "MatterOptions": _abi.MatterOptionsStruct, "AstroParams": _abi.AstroParamsStruct,
    "AstroOptions": _abi.AstroOptionsStruct, "CosmoTables": _abi.CosmoTablesStruct,
    "ConfigSettings": _abi.ConfigSettingsStruct, "InitialConditions": _abi.InitialConditionsStruct,
    "PerturbedField": _abi.PerturbedFieldStruct, "HaloBox": _abi.HaloBoxStruct, "TsBox": _abi.TsBoxStruct,
    "IonizedBox": _abi.IonizedBoxStruct, "BrightnessTemp": _abi.BrightnessTempStruct,
}


def _offsets_from_c(include_lines, tmp):
    """Compile a tiny C program printing sizeof/offsetof for every field of every struct."""
    body = []
    for name, st in STRUCTS.items():
        body.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f, _ in st._fields_:
            body.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    src = tmp / "abi.c"
    src.write_text("#include <stdio.h>\n#include <stddef.h>\n#include <stdbool.h>\n" + include_lines +
                   "\nint main(void){\n" + "\n".join(body) + "\nreturn 0;}\n")
    exe = tmp / "abi"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return dict(line.split() for line in out.splitlines())


def _offsets_from_ctypes():
    d = {}
    for name, st in STRUCTS.items():
        d[name] = str(C.sizeof(st))
        for f, _ in st._fields_:
            d[f"{name}.{f}"] = str(getattr(st, f).offset)
    return d


def test_header_layout_equals_ctypes_mirror(tmp_path):
    got = _offsets_from_c(f'#include "{HEADER}"', tmp_path)
    assert got == _offsets_from_ctypes()


def test_header_layout_equals_reference_headers(tmp_path):
    ref = Path("/root/reference/src/py21cmfast/src")
    if not ref.exists():
        pytest.skip("reference tree not present on this box")
    inc = f'#include "{ref}/_inputparams_wrapper.h"\n#include "{ref}/_outputstructs_wrapper.h"'
    got = _offsets_from_c(inc, tmp_path)
    assert got == _offsets_from_ctypes()


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    names = re.findall(r"^\s*(?:int|void|double|float)\s+\*?([A-Za-z_][A-Za-z0-9_]*)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    if not LIB.exists():
        pytest.skip("CUDA library not built (run __graft_entry__.build())")
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    fns = _declared_functions()
    assert len(fns) >= 25
    missing = [f for f in fns if f not in exported]
    assert not missing, missing
    for g in ("simulation_options_global", "matter_options_global", "cosmo_params_global", "astro_params_global",
              "astro_options_global", "cosmo_tables_global", "config_settings"):
        assert g in exported, g


def test_library_loads_without_gpu_and_product_has_no_fallback():
    """The library loads (dlopen) on a CPU-only box; the default backend path is the CUDA library
    and a missing library raises ImportError instead of falling back to anything on the CPU."""
    pkg = common.pkg
    if LIB.exists():
        be = pkg.Backend()
        assert Path(be.path).name == "lib21cmfast_b200.so"
    with pytest.raises(ImportError):
        pkg.Backend(ROOT / "21cmfast_b200" / "csrc" / "does_not_exist.so")
    src = (ROOT / "21cmfast_b200" / "_lib.py").read_text() + (ROOT / "21cmfast_b200" / "drivers.py").read_text()
    assert "oracle" not in src.replace("oracle/_ref", "").lower() or "import oracle" not in src
    for py in (ROOT / "21cmfast_b200").glob("*.py"):
        t = py.read_text()
        assert "from oracle" not in t and "import oracle" not in t and "_emu" not in t, py


def _reference_prototypes():
    ref = Path("/root/reference/src/py21cmfast/src/_functionprototypes_wrapper.h")
    if not ref.exists():
        pytest.skip("reference tree not present on this box")
    text = re.sub(r"/\*.*?\*/", "", ref.read_text(), flags=re.S)
    text = re.sub(r"//.*", "", text)
    return {re.search(r"(\w+)\s*\(", p).group(1): " ".join(p.split()) for p in text.split(";") if "(" in p}


def test_library_is_a_complete_link_target_for_the_reference_cffi_surface():
    """Every function and global the reference's cdef names (_functionprototypes_wrapper.h,
    _inputparams_wrapper.h) is exported, so the API-mode cffi build links against the library alone
    (SURVEY.md section 8b); the header declares the same prototypes."""
    if not LIB.exists():
        pytest.skip("CUDA library not built")
    protos = _reference_prototypes()
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    assert not [n for n in protos if n not in exported]
    assert "photon_cons_allocated" in exported
    declared = set(_declared_functions())
    assert not [n for n in protos if n not in declared], [n for n in protos if n not in declared]


def test_unscoped_symbols_fail_loudly(capfd):
    """Outside the scoped path nothing computes: status 3 / NaN and a message, never a silent CPU result."""
    import ctypes as C
    pkg = common.pkg
    if not LIB.exists():
        pytest.skip("CUDA library not built")
    lib = pkg.Backend().lib
    lib.InitialisePhotonCons.restype = C.c_int
    assert lib.InitialisePhotonCons() == 3
    lib.ComputeZstart_PhotonCons.restype = C.c_int
    assert lib.ComputeZstart_PhotonCons(None) == 3
    lib.expected_nhalo.restype = C.c_double
    lib.expected_nhalo.argtypes = [C.c_double]
    assert np.isnan(lib.expected_nhalo(8.0))
    lib.ComputeTau.restype = C.c_float
    assert np.isnan(lib.ComputeTau(0, None, None, C.c_float(3.0)))
    assert C.c_bool.in_dll(lib, "photon_cons_allocated").value is False
    err = capfd.readouterr().err
    assert "InitialisePhotonCons is outside the scoped hot path" in err and "expected_nhalo" in err


def test_get_sigma_matches_reference():
    import ctypes as C
    pkg = common.pkg
    emu, ref = common.emu_backend(), common.ref_backend()
    if emu is None or ref is None:
        pytest.skip("needs tests/_emu and oracle/_ref")
    inputs = common.make_inputs()
    masses = np.logspace(5.5, 15.5, 41)
    res = []
    for be in (emu, ref):
        be.state.init(inputs, broadcast_inputs=True, ps=True, sigma=True)
        s, d = np.zeros_like(masses), np.zeros_like(masses)
        fn = be.lib.get_sigma
        fn.restype = None
        fn.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        fn(len(masses), dp(masses), dp(s), dp(d))
        res.append((s, d))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-10)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-10)
    assert (np.diff(res[0][0]) < 0).all()
