"""Physics data table the scoped path needs at run time.

``init_heat`` (reference ``heating_helper_progs.c:94-197``) reads ``recfast_LCDM.dat`` through
``config_settings.external_table_path`` even when the spin temperature is off: ``T_RECFAST`` /
``xion_RECFAST`` give the no-T_s kinetic temperature and the neutral-box x_HI
(``IonisationBox.c:201-206,550``).  The reference ships the file in ``py21cmfast/_data``; this
package ships the same four columns as ``data/recfast_table.npz`` and writes the text file the C
side parses into a per-user cache directory on first use.  ``PY21CMFAST_DATA`` (a directory that
already holds ``recfast_LCDM.dat``, e.g. the reference's ``_data``) overrides it.
"""
from __future__ import annotations

import os
import tempfile
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_NPZ = _HERE / "data" / "recfast_table.npz"
_cached: Path | None = None


def write_recfast_table(directory: Path) -> Path:
    g = np.load(_NPZ)
    directory.mkdir(parents=True, exist_ok=True)
    target = directory / "recfast_LCDM.dat"
    tmp = directory / f".recfast_LCDM.{os.getpid()}.tmp"
    with open(tmp, "w") as f:
        for z, xe, c3, tk in zip(g["z"], g["xe"], g["col3"], g["tk"]):
            f.write(f"{z:8.2f}   {xe:.5E}    {c3:.5E}    {tk:.5E}\n")
    os.replace(tmp, target)  # atomic: several ranks may start at once
    return target


def default_table_dir() -> Path:
    """Directory to hand to ``config_settings.external_table_path``."""
    global _cached
    env = os.environ.get("PY21CMFAST_DATA")
    if env:
        if not Path(env, "recfast_LCDM.dat").exists():
            raise FileNotFoundError(f"PY21CMFAST_DATA={env} holds no recfast_LCDM.dat")
        return Path(env)
    if _cached is not None and (_cached / "recfast_LCDM.dat").exists():
        return _cached
    if not _NPZ.exists():
        raise FileNotFoundError(f"21cmfast_b200: packaged table {_NPZ} is missing")
    base = Path(os.environ.get("XDG_CACHE_HOME", Path.home() / ".cache")) / "21cmfast_b200"
    try:
        write_recfast_table(base)
    except OSError:
        base = Path(tempfile.mkdtemp(prefix="b200_tables_"))
        write_recfast_table(base)
    _cached = base
    return base
