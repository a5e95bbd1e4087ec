"""The C-ABI contract: struct layouts of include/py21cmfast_b200.h equal the ctypes mirror (and,
where /root/reference is present, the reference's own headers), and the built library exports
every function the header declares.  CPU only (no compute calls)."""
import ctypes as C
import re
import subprocess
import tempfile
from pathlib import Path

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
