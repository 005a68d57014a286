"""ctypes binding of libmapox_b200.so (the C ABI declared in include/mapox_b200.h).

There is no fallback: if the shared library is missing the import fails loudly, and every compute call
raises if the CUDA call fails (e.g. no sm_100 device).
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmapox_b200.so"


class MpxError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {_PKG / 'csrc'}` (nvcc, sm_100a). There is no CPU fallback."
        )
    return C.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else C.DEFAULT_MODE)


lib = _load()

c_ll = C.c_longlong
c_p = C.c_void_p


def _struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


LinearArgs = _struct("mpx_linear_args", [
    ("x", c_p), ("ldx", c_ll), ("wt", c_p), ("ldw", c_ll), ("bias", c_p), ("residual", c_p), ("ldr", c_ll),
    ("y", c_p), ("ldy", c_ll), ("y_f32", c_p), ("ldy32", c_ll), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
    ("act", C.c_int)])

WgradArgs = _struct("mpx_wgrad_args", [
    ("x", c_p), ("ldx", c_ll), ("dy", c_p), ("ldy", c_ll), ("dw", c_p), ("lddw", c_ll),
    ("M", C.c_int), ("N", C.c_int), ("K", C.c_int)])

QkvRopeArgs = _struct("mpx_qkv_rope_args", [
    ("x", c_p), ("ldx", c_ll), ("wt_qkv", c_p), ("pos", c_p), ("pos_len", C.c_int), ("rope_timescale", c_p),
    ("q", c_p), ("k", c_p), ("k_row_stride", c_ll), ("v", c_p), ("v_row_stride", c_ll), ("cache_S", C.c_int),
    ("M", C.c_int), ("H", C.c_int), ("Nq", C.c_int), ("Nkv", C.c_int), ("D", C.c_int)])

GluUpArgs = _struct("mpx_glu_up_args", [
    ("x", c_p), ("ldx", c_ll), ("wt_ug", c_p), ("h", c_p), ("u_save", c_p), ("g_save", c_p),
    ("M", C.c_int), ("F", C.c_int), ("K", C.c_int)])

AttnDecodeArgs = _struct("mpx_attn_decode_args", [
    ("q", c_p), ("kcache", c_p), ("vcache", c_p), ("pos", c_p), ("out", c_p),
    ("B", C.c_int), ("S", C.c_int), ("Nq", C.c_int), ("Nkv", C.c_int), ("D", C.c_int)])

lib.mpx_last_error.restype = C.c_char_p
lib.mpx_gae_workspace_bytes.restype = C.c_size_t
lib.mpx_gae_workspace_bytes.argtypes = [C.c_int, C.c_int]
lib.mpx_gae.argtypes = [c_p, c_p, c_p, C.c_int, C.c_int, C.c_double, C.c_double, c_p, c_p, c_p, c_p, c_p]
lib.mpx_adv_normalize.argtypes = [c_p, c_ll, c_p, C.c_double, c_p]
lib.mpx_hlgauss_value.argtypes = [c_p, c_ll, C.c_int, c_p, c_p, c_p]
lib.mpx_hlgauss_loss_workspace_bytes.restype = C.c_size_t
lib.mpx_hlgauss_loss_workspace_bytes.argtypes = [c_ll]
lib.mpx_hlgauss_loss.argtypes = [c_p, c_p, c_ll, C.c_int, c_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                 c_p, c_p, c_p, c_p]
for _name in ("mpx_linear", "mpx_linear_wgrad", "mpx_qkv_rope", "mpx_glu_up", "mpx_attn_decode"):
    getattr(lib, _name).argtypes = [c_p, c_p]
lib.mpx_debug_set.argtypes = [C.c_int, C.c_int]
lib.mpx_rmsnorm_fwd.argtypes = [c_p, c_p, C.c_float, c_p, c_ll, C.c_int, c_p]
lib.mpx_rmsnorm_bwd.argtypes = [c_p, c_p, c_p, C.c_float, c_p, c_p, c_p, c_ll, C.c_int, c_p]
lib.mpx_glu_bwd.argtypes = [c_p, c_p, c_p, c_p, c_ll, C.c_int, c_p]
lib.mpx_gelu_bwd.argtypes = [c_p, c_p, c_p, c_ll, c_p]
lib.mpx_rope.argtypes = [c_p, c_ll, c_p, C.c_int, c_p, c_ll, C.c_int, C.c_int, C.c_int, c_p]
lib.mpx_colsum.argtypes = [c_p, c_ll, c_ll, C.c_int, c_p, c_p]
lib.mpx_add_bf16.argtypes = [c_p, c_p, c_p, c_ll, c_p]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.mpx_last_error()
        raise MpxError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
