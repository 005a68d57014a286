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


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every ctypes mirror in jaxrl_b200/_lib.py has the size and field offsets gcc gives the struct of the same name
    in include/mapox_b200.h (catches a field added on one side only)."""
    import shutil
    import subprocess

    import pytest
    from jaxrl_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = [v for v in vars(_lib).values()
               if isinstance(v, type) and issubclass(v, ctypes.Structure) and v is not ctypes.Structure
               and v.__name__.startswith("mpx_")]
    assert len(structs) >= 8
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mapox_b200.h"', 'int main(void) {']
    for st in structs:
        lines.append(f'  printf("{st.__name__} %zu\\n", sizeof({st.__name__}));')
        for name, _ in st._fields_:
            lines.append(f'  printf("{st.__name__}.{name} %zu\\n", offsetof({st.__name__}, {name}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for st in structs:
        assert int(out[st.__name__]) == ctypes.sizeof(st), f"sizeof({st.__name__})"
        for name, _ in st._fields_:
            assert int(out[f"{st.__name__}.{name}"]) == getattr(st, name).offset, f"offsetof({st.__name__}, {name})"
