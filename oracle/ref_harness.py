"""Bind the compiled reference (oracle/_ref/libref21cmfast.so) -- TEST INFRASTRUCTURE ONLY.

The reference's C sources are compiled unmodified against the shim headers by ``make -C oracle
ref`` (only possible where /root/reference exists; the resulting .so and the small data tables it
reads travel to the GPU box inside the git-ignored ``oracle/_ref``).  This module exposes it
through the same ``Backend`` class the product uses, so a parity test is literally
"same inputs, two shared libraries".  Only tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke() may import this.
"""
from __future__ import annotations

import importlib
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_LIB = HERE / "_ref" / "libref21cmfast.so"
REF_DATA = HERE / "_ref" / "data"
REFERENCE_DATA = Path("/root/reference/src/py21cmfast/_data")


def _torch_cpu_lib():
    try:
        import torch
        p = Path(torch.__file__).parent / "lib" / "libtorch_cpu.so"
        return str(p) if p.exists() else None
    except Exception:
        return None


def table_path() -> Path | None:
    for p in (os.environ.get("PY21CMFAST_DATA"), REF_DATA, REFERENCE_DATA):
        if p and Path(p, "recfast_LCDM.dat").exists():
            return Path(p)
    return None


def available() -> bool:
    return REF_LIB.exists()


_backend = None


def ref_backend(fft: str = "mkl", fft_threads: int = 1):
    """Backend bound to the compiled reference.  ``fft`` selects the FFTW-shim back-end."""
    global _backend
    os.environ["ORACLE_FFT_THREADS"] = str(fft_threads)
    if _backend is None:
        os.environ.setdefault("ORACLE_FFT", fft)
        lib = _torch_cpu_lib()
        if lib and fft == "mkl":
            os.environ.setdefault("ORACLE_TORCH_LIB", lib)
        pkg = importlib.import_module("21cmfast_b200")
        _backend = pkg.Backend(REF_LIB)
        tp = table_path()
        if tp is not None:
            _backend.set_table_path(tp)
        _backend.set_wisdoms_path("/tmp")
    return _backend
