"""ctypes binding of libmapox_b200.so (the C ABI declared in include/mapox_b200.h).

There is no fallback: if the shared library is missing the import fails loudly, and every compute call
raises if the CUDA call fails (e.g. no sm_100 device).
"""

from __future__ import annotations

import ctypes as C
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
    return C.CDLL(str(LIB_PATH))


lib = _load()

c_ll = C.c_longlong
c_p = C.c_void_p
c_i = C.c_int
c_f = C.c_float


def _struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


LinearArgs = _struct("mpx_linear_args", [
    ("x", c_p), ("ldx", c_ll), ("wt", c_p), ("ldw", c_ll), ("bias", c_p), ("residual", c_p), ("ldr", c_ll),
    ("y", c_p), ("ldy", c_ll), ("y_f32", c_p), ("ldy32", c_ll), ("M", c_i), ("N", c_i), ("K", c_i),
    ("act", c_i), ("y_pre", c_p)])

WgradArgs = _struct("mpx_wgrad_args", [
    ("x", c_p), ("ldx", c_ll), ("dy", c_p), ("ldy", c_ll), ("dw", c_p), ("lddw", c_ll),
    ("M", c_i), ("N", c_i), ("K", c_i), ("workspace", c_p), ("workspace_bytes", C.c_size_t)])

QkvRopeArgs = _struct("mpx_qkv_rope_args", [
    ("x", c_p), ("ldx", c_ll), ("wt_qkv", c_p), ("pos", c_p), ("pos_len", c_i), ("rope_timescale", c_p),
    ("q", c_p), ("q_row_stride", c_ll), ("k", c_p), ("k_row_stride", c_ll), ("v", c_p), ("v_row_stride", c_ll),
    ("cache_S", c_i), ("M", c_i), ("H", c_i), ("Nq", c_i), ("Nkv", c_i), ("D", c_i),
    ("norm_scale", c_p), ("norm_eps", C.c_float)])

GluUpArgs = _struct("mpx_glu_up_args", [
    ("x", c_p), ("ldx", c_ll), ("wt_ug", c_p), ("h", c_p), ("u_save", c_p), ("g_save", c_p),
    ("M", c_i), ("F", c_i), ("K", c_i)])

AttnDecodeArgs = _struct("mpx_attn_decode_args", [
    ("q", c_p), ("kcache", c_p), ("vcache", c_p), ("pos", c_p), ("out", c_p),
    ("B", c_i), ("S", c_i), ("Nq", c_i), ("Nkv", c_i), ("D", c_i)])

AttnDecodeFusedArgs = _struct("mpx_attn_decode_fused_args", [
    ("x", c_p), ("norm_scale", c_p), ("eps", C.c_float), ("wt_qkv", c_p), ("pos", c_p), ("rope_timescale", c_p),
    ("kcache", c_p), ("vcache", c_p), ("out", c_p),
    ("B", c_i), ("S", c_i), ("H", c_i), ("Nq", c_i), ("Nkv", c_i), ("D", c_i), ("flags", c_i)])

AttnTrainArgs = _struct("mpx_attn_train_args", [
    ("q", c_p), ("ldq", c_ll), ("k", c_p), ("ldk", c_ll), ("v", c_p), ("ldv", c_ll), ("o", c_p), ("ldo", c_ll),
    ("lse", c_p), ("B", c_i), ("T", c_i), ("Nq", c_i), ("Nkv", c_i), ("D", c_i)])

AttnTrainBwdArgs = _struct("mpx_attn_train_bwd_args", [
    ("q", c_p), ("ldq", c_ll), ("k", c_p), ("ldk", c_ll), ("v", c_p), ("ldv", c_ll), ("o", c_p), ("ldo", c_ll),
    ("dout", c_p), ("lddo", c_ll), ("lse", c_p), ("delta", c_p), ("dq_acc", c_p),
    ("dq", c_p), ("lddq", c_ll), ("dk", c_p), ("lddk", c_ll), ("dv", c_p), ("lddv", c_ll),
    ("pos", c_p), ("pos_len", c_i), ("rope_timescale", c_p),
    ("B", c_i), ("T", c_i), ("Nq", c_i), ("Nkv", c_i), ("D", c_i)])

PolicyHeadArgs = _struct("mpx_policy_head_args", [
    ("norm_scale", c_p), ("eps", C.c_float), ("table", c_p), ("A", c_i), ("action_mask", c_p), ("gumbel_in", c_p),
    ("seed", C.c_ulonglong), ("stream_id", C.c_ulonglong), ("stream_offset", c_p), ("logits_out", c_p), ("xf_out", c_p),
    ("action", c_p), ("log_prob", c_p)])

PrepDesc = _struct("mpx_prep_desc", [
    ("src", c_p), ("src_ld", c_ll), ("dst", c_p), ("dst_ld", c_ll), ("rows", c_i), ("cols", c_i),
    ("transpose", c_i), ("_pad", c_i)])

lib.mpx_last_error.restype = C.c_char_p
lib.mpx_debug_set.argtypes = [c_i, c_i]
lib.mpx_debug_set_stamps.argtypes = [c_p]
lib.mpx_gae_workspace_bytes.restype = C.c_size_t
lib.mpx_gae_workspace_bytes.argtypes = [c_i, c_i]
lib.mpx_gae.argtypes = [c_p, c_p, c_p, c_i, c_i, C.c_double, C.c_double, c_p, c_p, c_p, c_p, c_p]
lib.mpx_adv_normalize.argtypes = [c_p, c_ll, c_p, C.c_double, c_p]
lib.mpx_hlgauss_value.argtypes = [c_p, c_ll, c_i, c_p, c_p, c_p]
lib.mpx_hlgauss_loss_workspace_bytes.restype = C.c_size_t
lib.mpx_hlgauss_loss_workspace_bytes.argtypes = [c_ll]
lib.mpx_hlgauss_loss.argtypes = [c_p, c_p, c_ll, c_i, c_p, c_f, c_f, c_f, c_f, c_p, c_p, c_p, c_i, c_p, c_p]
for _name in ("mpx_linear", "mpx_linear_wgrad", "mpx_qkv_rope", "mpx_glu_up", "mpx_attn_decode",
              "mpx_attn_decode_fused", "mpx_attn_train_fwd", "mpx_attn_train_bwd"):
    getattr(lib, _name).argtypes = [c_p, c_p]
lib.mpx_linear_wgrad_workspace_bytes.restype = C.c_size_t
lib.mpx_linear_wgrad_workspace_bytes.argtypes = [c_i, c_i, c_i]
lib.mpx_linear_f32w.argtypes = [c_p, c_ll, c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_p]
lib.mpx_glu_fused.argtypes = [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, C.c_float, c_p, c_p, c_p]
lib.mpx_glu_fused_head.argtypes = [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, C.c_float, c_p, C.POINTER(PolicyHeadArgs), c_p]
lib.mpx_glu_bwd_fused.argtypes = [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p]
lib.mpx_value_head_fused.argtypes = [c_p, c_p, C.c_float, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]
lib.mpx_value_bwd_fused_workspace_bytes.restype = C.c_size_t
lib.mpx_value_bwd_fused_workspace_bytes.argtypes = [c_i]
lib.mpx_value_bwd_fused.argtypes = [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, C.c_size_t, c_p]
lib.mpx_rmsnorm_fwd.argtypes = [c_p, c_p, c_f, c_p, c_ll, c_i, c_p]
lib.mpx_rmsnorm_bwd.argtypes = [c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_ll, c_i, c_p]
lib.mpx_glu_bwd.argtypes = [c_p, c_p, c_p, c_p, c_ll, c_i, c_p]
lib.mpx_gelu_bwd.argtypes = [c_p, c_p, c_p, c_ll, c_p]
lib.mpx_rope.argtypes = [c_p, c_ll, c_p, c_i, c_p, c_ll, c_i, c_i, c_i, c_p]
lib.mpx_colsum.argtypes = [c_p, c_ll, c_ll, c_i, c_p, c_p]
lib.mpx_add_bf16.argtypes = [c_p, c_p, c_p, c_ll, c_p]
lib.mpx_embed_combine.argtypes = [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_p, c_ll, c_i, c_p]
lib.mpx_embed_bwd.argtypes = [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_ll, c_i, c_p]
lib.mpx_policy_logits.argtypes = [c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_p]
lib.mpx_policy_sample.argtypes = [c_p, c_p, c_p, c_p, c_ll, c_i, c_p]
lib.mpx_gumbel_noise.argtypes = [c_p, c_ll, C.c_ulonglong, C.c_ulonglong, c_p, c_p]
lib.mpx_conv1_onehot_wgrad_workspace_bytes.restype = C.c_size_t
lib.mpx_conv1_onehot_wgrad_workspace_bytes.argtypes = [c_ll, c_i, c_i, c_i, c_i, c_i, c_i]
lib.mpx_conv1_onehot_wgrad.argtypes = [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p, C.c_size_t, c_p]
lib.mpx_flat_embed_fwd.argtypes = [c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_p]
lib.mpx_flat_embed_bwd.argtypes = [c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_p]
lib.mpx_obs_fused_image_bytes.restype = C.c_size_t
lib.mpx_obs_fused_image_bytes.argtypes = []
lib.mpx_obs_fused_prep.argtypes = [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]
lib.mpx_obs_embed_fused.argtypes = [c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_p, c_p, c_ll, c_i, c_i,
                                    c_i, c_p]
lib.mpx_obs_train_supported.argtypes = [c_p, c_i]
lib.mpx_obs_train_tiles.restype = c_ll
lib.mpx_obs_train_tiles.argtypes = [c_ll]
lib.mpx_obs_train_fwd.argtypes = [c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p,
                                  c_ll, c_i, c_i, c_i, c_p]
lib.mpx_obs_train_bwd_workspace_bytes.restype = C.c_size_t
lib.mpx_obs_train_bwd_workspace_bytes.argtypes = []
lib.mpx_obs_train_bwd.argtypes = [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_p, C.c_size_t, c_p]
lib.mpx_obs_conv2_wgrad.argtypes = [c_p, c_p, c_p, c_ll, c_p, C.c_size_t, c_p]
lib.mpx_policy_head_sample.argtypes = [c_p, c_p, C.c_float, c_p, c_p, c_p, C.c_ulonglong, C.c_ulonglong, c_p, c_p, c_p, c_p,
                                       c_p, c_ll, c_i, c_i, c_p]
lib.mpx_ppo_policy_loss_workspace_bytes.restype = C.c_size_t
lib.mpx_ppo_policy_loss_workspace_bytes.argtypes = [c_ll]
lib.mpx_ppo_policy_loss.argtypes = [c_p, c_p, c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_p, c_ll, c_i, c_p]
lib.mpx_policy_head_bwd.argtypes = [c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_p]
lib.mpx_obs_onehot_im2col.argtypes = [c_p, c_p, c_ll, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p]
lib.mpx_conv1_onehot_fwd.argtypes = [c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p]
lib.mpx_im2col.argtypes = [c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p]
lib.mpx_col2im.argtypes = [c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]
lib.mpx_adamw_workspace_bytes.restype = C.c_size_t
lib.mpx_adamw_workspace_bytes.argtypes = [c_ll]
lib.mpx_adamw_step.argtypes = [c_p, c_p, c_p, c_p, c_p, c_ll, c_p, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_p,
                               c_p, c_p]
lib.mpx_muon_workspace_bytes.restype = C.c_size_t
lib.mpx_muon_workspace_bytes.argtypes = [c_ll, c_i]
lib.mpx_muon_step.argtypes = [c_p, c_p, c_p, c_p, c_p, c_ll, c_ll, c_p, c_i, c_ll, c_ll, c_i, c_i, c_p, c_p, c_f, c_f, c_f,
                              c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_p, c_p, c_p]
lib.mpx_muon_step_phased.argtypes = [c_p, c_p, c_p, c_p, c_p, c_ll, c_ll, c_p, c_i, c_ll, c_ll, c_i, c_i, c_p, c_p, c_f, c_f,
                                     c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_p, c_p, c_i, c_p]
lib.mpx_muon_apply_sharded.argtypes = [c_p, c_p, c_p, c_p, c_p, c_ll, c_ll, c_p, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_p,
                                       c_p, c_p]
lib.mpx_prep_weights.argtypes = [c_p, c_i, c_p]
lib.mpx_transpose_tb.argtypes = [c_p, c_p, c_p, c_i, c_i, c_i, c_p]


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
