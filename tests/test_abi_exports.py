"""The C-ABI library loads on a CPU-only box and exports every symbol include/mapox_b200.h declares."""

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "mapox_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpx_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "mpx_gae" in syms and "mpx_attn_decode" in syms and len(syms) >= 10


def test_library_exports_every_declared_symbol():
    from jaxrl_b200 import _lib
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_no_cpu_fallback_errors_are_loud():
    """Argument validation works without a GPU and reports through mpx_last_error()."""
    from jaxrl_b200 import _lib
    rc = _lib.lib.mpx_gae(None, None, None, 0, 0, 0.9, 0.9, None, None, None, None, None)
    assert rc != 0 and b"mpx_gae" in _lib.lib.mpx_last_error()
    assert _lib.lib.mpx_version() >= 100
