"""JAX side of the drop-in boundary: registers the XLA-FFI handlers of libmapox_b200_xla.so (csrc/xla_ffi/xla_ffi.cc)
as jax.ffi custom-call targets and wraps the training ops in jax.custom_vjp, so that mapox_trainer's JAX host code
(model/attention.py, model/feed_forward.py, rollout.py, values.py, train.py) calls the sm_100a kernels instead of
XLA/cuDNN.  INTEGRATION.md shows the reference-side patch that uses these functions.

UNVERIFIED HERE: jax / jaxlib are not installable in this image, so this module cannot even be
imported in the build container; `tests/test_xla_ffi_source.py` checks it statically (syntax, one registration per
handler symbol, operand / result / attribute counts of every ffi_call against the C++ Bind() chain).  There is no
fallback: importing it without jax or without the shared library raises.

Layout conventions match the reference's arrays: activations (B,T,H) bf16, KV rings (B,S,Nkv,D) bf16,
positions int32, flax kernels as stored.  Weight operand layouts the kernels want (transposed / packed bf16 copies) are
produced with jnp ops in `stage_*` below - XLA fuses them and they run once per weight update.
"""

from __future__ import annotations

import ctypes
import functools
import math
import os

import jax
import jax.numpy as jnp
import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_XLA_LIB = ctypes.CDLL(os.path.join(_DIR, "libmapox_b200_xla.so"))      # raises if missing: no fallback
_CORE = ctypes.CDLL(os.path.join(_DIR, "libmapox_b200.so"))
for _n, _args in (("mpx_gae_workspace_bytes", [ctypes.c_int, ctypes.c_int]),
                  ("mpx_hlgauss_loss_workspace_bytes", [ctypes.c_longlong]),
                  ("mpx_ppo_policy_loss_workspace_bytes", [ctypes.c_longlong]),
                  ("mpx_linear_wgrad_workspace_bytes", [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
                  ("mpx_conv1_onehot_wgrad_workspace_bytes", [ctypes.c_longlong] + [ctypes.c_int] * 6),
                  ("mpx_adamw_workspace_bytes", [ctypes.c_longlong]),
                  ("mpx_muon_workspace_bytes", [ctypes.c_longlong, ctypes.c_int]),
                  ("mpx_obs_fused_image_bytes", []), ("mpx_obs_train_bwd_workspace_bytes", []),
                  ("mpx_value_bwd_fused_workspace_bytes", [ctypes.c_int])):
    getattr(_CORE, _n).restype = ctypes.c_size_t
    getattr(_CORE, _n).argtypes = _args

# one entry per XLA_FFI_DEFINE_HANDLER_SYMBOL in csrc/xla_ffi/xla_ffi.cc
TARGETS = (
    "MpxGae", "MpxAdvNormalize", "MpxHlGaussValue", "MpxHlGaussLoss", "MpxLinear", "MpxLinearWgrad", "MpxLinearF32w",
    "MpxQkvRopeDecode", "MpxQkvRopeTrain", "MpxRope", "MpxAttnDecode", "MpxAttnBlockDecode", "MpxAttnTrainFwd",
    "MpxAttnTrainBwd", "MpxGluUp", "MpxGluFused", "MpxGluFusedHead", "MpxGluBwdFused", "MpxGluBwd", "MpxRmsNormFwd", "MpxRmsNormBwd",
    "MpxGeluBwd", "MpxColSum", "MpxAddBf16", "MpxEmbedCombine", "MpxEmbedBwd", "MpxObsFusedPrep", "MpxObsEmbedFused",
    "MpxConv1OnehotFwd", "MpxConv1OnehotWgrad", "MpxObsOnehotIm2col", "MpxIm2col", "MpxCol2im", "MpxFlatEmbedFwd",
    "MpxFlatEmbedBwd", "MpxPolicyLogits", "MpxPolicySample", "MpxPolicyHeadSample", "MpxGumbelNoise", "MpxPpoPolicyLoss",
    "MpxPolicyHeadBwd", "MpxValueHeadFused", "MpxAdamwStep", "MpxMuonStep", "MpxMuonApplySharded", "MpxPrepWeights",
    "MpxTransposeTb", "MpxObsTrainFwd", "MpxObsTrainBwd", "MpxObsConv2Wgrad", "MpxValueBwdFused",
)
for _t in TARGETS:
    jax.ffi.register_ffi_target(_t, jax.ffi.pycapsule(getattr(_XLA_LIB, _t)), platform="CUDA")

BF16, F32, I32, U8 = jnp.bfloat16, jnp.float32, jnp.int32, jnp.uint8
_EMPTY_BF = functools.partial(jnp.zeros, (0,), BF16)
_EMPTY_F32 = functools.partial(jnp.zeros, (0,), F32)


def _sds(shape, dtype):
    return jax.ShapeDtypeStruct(tuple(shape), dtype)


def _call(target, results, *operands, aliases=None, **attrs):
    return jax.ffi.ffi_call(target, results, input_output_aliases=aliases or {}, vmap_method="broadcast_all")(*operands, **attrs)


# ------------------------------------------------------------------------------------------------ weight staging
def stage_linear(kernel, out_dims: int = 1):
    """flax kernel (..in.., ..out..) f32 -> (wt, w): the transposed (N,K) bf16 operand of the forward GEMM and the
    as-stored (K,N) bf16 operand of the data gradient."""
    k = kernel.reshape(-1, math.prod(kernel.shape[-out_dims:])) if out_dims else kernel
    return k.T.astype(BF16), k.astype(BF16)


def stage_qkv(wq, wk, wv):
    """query/key/value LinearGeneral kernels (H,N,D) -> wt_qkv ((Nq+2Nkv)*D, H) bf16 and w_qkv (H, (Nq+2Nkv)*D)."""
    H = wq.shape[0]
    w = jnp.concatenate([wq.reshape(H, -1), wk.reshape(H, -1), wv.reshape(H, -1)], axis=1)
    return w.T.astype(BF16), w.astype(BF16)


def stage_glu64(w_up, w_gate):
    """(H,F) up / gate kernels -> (2F,H) bf16 in 128-row tiles [64 up rows | 64 gate rows] (mpx_glu_fused)."""
    H, F = w_up.shape
    ut = w_up.T.reshape(F // 64, 64, H)
    gt = w_gate.T.reshape(F // 64, 64, H)
    return jnp.concatenate([ut, gt], axis=1).reshape(2 * F, H).astype(BF16)


def rope_timescale(head_dim: int, max_wavelength: float = 10_000.0):
    return jnp.asarray(max_wavelength, F32) ** (2.0 * jnp.arange(0, head_dim // 2, dtype=F32) / head_dim)


# ------------------------------------------------------------------------------------------------ GAE / HL-Gauss
def gae(rewards, values, next_terminated, discount: float, gae_lambda: float):
    """Rollout.calculate_advantage before normalisation (rollout.py:128-157) -> (advantages, targets, stats f64[2])."""
    B, T = rewards.shape
    adv, tgt, stats, _ = _call("MpxGae", (_sds((B, T), F32), _sds((B, T), F32), _sds((2,), jnp.float64),
                                          _sds((_CORE.mpx_gae_workspace_bytes(B, T),), U8)),
                               rewards, values, next_terminated, discount=np.float64(discount), gae_lambda=np.float64(gae_lambda))
    return adv, tgt, stats


def adv_normalize(advantages, stats, count: float):
    return _call("MpxAdvNormalize", _sds(advantages.shape, F32), advantages, stats, aliases={0: 0}, count=np.float64(count))


def hlgauss_value(logits, centers):
    """HlGaussValue.get_value (values.py:43-45)."""
    return _call("MpxHlGaussValue", _sds(logits.shape[:-1], F32), logits, centers)


@functools.partial(jax.custom_vjp, nondiff_argnums=(2, 3, 4, 5))
def hlgauss_loss(logits, targets, supports, vmin, vmax, sigma):
    """HlGaussValue.get_loss (values.py:47-63); the gradient w.r.t. logits comes from the same kernel pass."""
    return _hlgauss_loss_fwd(logits, targets, supports, vmin, vmax, sigma)[0]


def _hlgauss_loss_fwd(logits, targets, supports, vmin, vmax, sigma):
    n = math.prod(logits.shape[:-1])
    loss, dlogits, _, _ = _call("MpxHlGaussLoss", (_sds((1,), F32), _sds(logits.shape, F32), _sds((0,), BF16),
                                                   _sds((_CORE.mpx_hlgauss_loss_workspace_bytes(n),), U8)),
                                logits, targets, supports, vmin=np.float32(vmin), vmax=np.float32(vmax),
                                sigma=np.float32(sigma), grad_scale=np.float32(1.0))
    return loss[0], dlogits


def _hlgauss_loss_bwd(supports, vmin, vmax, sigma, dlogits, g):
    return g * dlogits, None


hlgauss_loss.defvjp(_hlgauss_loss_fwd, _hlgauss_loss_bwd)


# ------------------------------------------------------------------------------------------------ linear (custom_vjp)
def _linear_raw(x, wt, bias=None, residual=None, act: int = 0, save_pre: bool = False):
    M = math.prod(x.shape[:-1])
    N = wt.shape[0]
    y, y_pre = _call("MpxLinear", (_sds(x.shape[:-1] + (N,), BF16), _sds((M, N) if save_pre else (0,), BF16)),
                     x, wt, _EMPTY_BF() if bias is None else bias, _EMPTY_BF() if residual is None else residual,
                     act=np.int32(act))
    return y, y_pre


def _wgrad_raw(x, dy):
    M, K, N = math.prod(x.shape[:-1]), x.shape[-1], dy.shape[-1]
    dw, _ = _call("MpxLinearWgrad", (_sds((K, N), F32), _sds((_CORE.mpx_linear_wgrad_workspace_bytes(M, N, K),), U8)),
                  x, dy, jnp.zeros((K, N), F32), aliases={2: 0})
    return dw


@jax.custom_vjp
def linear(x, kernel):
    """nnx.Linear / LinearGeneral without bias, dtype=bf16 (attention.py:53-92, feed_forward.py:59-71): x (...,K) bf16,
    kernel (K,N) f32 master parameter -> (...,N) bf16.  Forward, data gradient and weight gradient are tcgen05 GEMMs."""
    return _linear_raw(x, kernel.T.astype(BF16))[0]


def _linear_fwd(x, kernel):
    return _linear_raw(x, kernel.T.astype(BF16))[0], (x, kernel)


def _linear_bwd(res, dy):
    x, kernel = res
    dx = _linear_raw(dy, kernel.astype(BF16))[0]              # dx = dy @ W^T: the kernel as stored is the (N_out=K, K_in=N) operand
    return dx, _wgrad_raw(x, dy)


linear.defvjp(_linear_fwd, _linear_bwd)


# ------------------------------------------------------------------------------------------------ attention
def attn_block_decode(x, norm_scale, wt_qkv, seq_pos, timescale, kcache, vcache, num_heads: int, eps: float = 1e-6):
    """history_norm + q/k/v + RoPE + update_kv_cache + decode SDPA of one rollout step (network.py:158-165,
    attention.py:96-150) -> (attention output (B, Nq*D), kcache, vcache); the rings are updated in place (aliased)."""
    B = kcache.shape[0]
    D = kcache.shape[3]
    return _call("MpxAttnBlockDecode", (_sds((B, num_heads * D), BF16), _sds(kcache.shape, BF16), _sds(vcache.shape, BF16)),
                 x, norm_scale, wt_qkv, seq_pos, timescale, kcache, vcache, aliases={5: 1, 6: 2},
                 num_heads=np.int32(num_heads), eps=np.float32(eps))


def qkv_rope_decode(x, wt_qkv, seq_pos, timescale, kcache, vcache, num_heads: int, norm_scale=None, eps: float = 1e-6):
    """q/k/v projection + RoPE + ring write (attention.py:121-134, 108-114) for shapes without the fused block kernel."""
    B, D = kcache.shape[0], kcache.shape[3]
    return _call("MpxQkvRopeDecode", (_sds((B, num_heads * D), BF16), _sds(kcache.shape, BF16), _sds(vcache.shape, BF16)),
                 x, wt_qkv, seq_pos, timescale, kcache, vcache, _EMPTY_F32() if norm_scale is None else norm_scale,
                 aliases={4: 1, 5: 2}, num_heads=np.int32(num_heads), eps=np.float32(eps))


def attn_decode(q, kcache, vcache, seq_pos):
    """Decode-branch SDPA (attention.py:136-150): q (B,1,Nq,D) against the rings, kv_len = min(seq_pos[0,0]+1, S)."""
    return _call("MpxAttnDecode", _sds(q.shape, BF16), q, kcache, vcache, seq_pos)


@functools.partial(jax.custom_vjp, nondiff_argnums=(3, 4, 5))
def attn_train(qkv_unrotated_dummy, qkv, positions, num_heads, num_kv_heads, head_dim):
    """Causal GQA attention over a full window (attention.py:151-169) on the packed, already rotated qkv
    (B,T,(Nq+2Nkv)*D) produced by `qkv_rope_train`.  `qkv_unrotated_dummy` is the differentiable handle: the backward
    kernel returns the cotangents of the UN-rotated projections (inverse RoPE fused), so the custom_vjp is defined from
    the un-rotated projection output; see `attention_train` below for the composed function the model calls."""
    del qkv_unrotated_dummy
    return _attn_train_fwd_call(qkv, num_heads, num_kv_heads, head_dim)[0]


def _attn_train_fwd_call(qkv, num_heads, num_kv_heads, head_dim):
    B, T = qkv.shape[0], qkv.shape[1]
    return _call("MpxAttnTrainFwd", (_sds((B, T, num_heads * head_dim), BF16), _sds((B, T, num_heads), F32)), qkv,
                 num_heads=np.int32(num_heads), num_kv_heads=np.int32(num_kv_heads), head_dim=np.int32(head_dim))


def _attn_train_fwd(qkv_unrotated_dummy, qkv, positions, num_heads, num_kv_heads, head_dim):
    o, lse = _attn_train_fwd_call(qkv, num_heads, num_kv_heads, head_dim)
    return o, (qkv, o, lse, positions)


def _attn_train_bwd(num_heads, num_kv_heads, head_dim, res, do):
    qkv, o, lse, positions = res
    B, T = qkv.shape[0], qkv.shape[1]
    dqkv, _, _ = _call("MpxAttnTrainBwd", (_sds(qkv.shape, BF16), _sds((B, T, num_heads), F32),
                                           _sds((B, T, num_heads * head_dim), F32)),
                       qkv, o, do, lse, positions, rope_timescale(head_dim),
                       num_heads=np.int32(num_heads), num_kv_heads=np.int32(num_kv_heads), head_dim=np.int32(head_dim))
    return dqkv, jnp.zeros_like(qkv), None


attn_train.defvjp(_attn_train_fwd, _attn_train_bwd)


def qkv_rope_train(xn, wt_qkv, positions, num_heads: int, num_kv_heads: int, head_dim: int):
    B, T = xn.shape[0], xn.shape[1]
    return _call("MpxQkvRopeTrain", _sds((B, T, (num_heads + 2 * num_kv_heads) * head_dim), BF16), xn, wt_qkv, positions,
                 rope_timescale(head_dim), num_heads=np.int32(num_heads), num_kv_heads=np.int32(num_kv_heads),
                 head_dim=np.int32(head_dim))


def attention_train(xn, wq, wk, wv, positions, num_heads: int, num_kv_heads: int, head_dim: int):
    """q/k/v projection + RoPE + causal attention of the training branch (attention.py:121-134,151-169), differentiable
    w.r.t. xn and the three kernels: the un-rotated projection is a `linear` (its own custom_vjp), the rotation and the
    attention are one custom_vjp whose backward returns un-rotated cotangents."""
    H = xn.shape[-1]
    w = jnp.concatenate([wq.reshape(H, -1), wk.reshape(H, -1), wv.reshape(H, -1)], axis=1)
    proj = linear(xn, w)                                       # un-rotated q|k|v, differentiable
    rotated = jax.lax.stop_gradient(qkv_rope_train(xn, w.T.astype(BF16), positions, num_heads, num_kv_heads, head_dim))
    return attn_train(proj, rotated, positions, num_heads, num_kv_heads, head_dim)


# ------------------------------------------------------------------------------------------------ GLU (custom_vjp)
@jax.custom_vjp
def glu_block(x1, norm_scale, w_up, w_gate, w_down):
    """x1 + GLUBlock(RMSNorm(x1)) (network.py:167-170, feed_forward.py:73-78), H = 128: one kernel forward (h kept for
    dWdown), one kernel backward (u, g, dh recomputed on chip)."""
    return _glu_fwd(x1, norm_scale, w_up, w_gate, w_down)[0]


def _glu_fwd(x1, norm_scale, w_up, w_gate, w_down):
    F = w_up.shape[1]
    M = math.prod(x1.shape[:-1])
    xn = _call("MpxRmsNormFwd", _sds(x1.shape, BF16), x1, norm_scale, eps=np.float32(1e-6))
    wt_ug64, wt_d = stage_glu64(w_up, w_gate), w_down.T.astype(BF16)
    out, h = _call("MpxGluFused", (_sds(x1.shape, BF16), _sds((M, F), BF16)), xn, wt_ug64, wt_d, x1, _EMPTY_F32(), _EMPTY_BF(),
                   eps=np.float32(1e-6))
    return out, (x1, xn, h, norm_scale, wt_ug64, wt_d)


def _glu_bwd(res, dy):
    x1, xn, h, norm_scale, wt_ug64, wt_d = res
    M, H = math.prod(x1.shape[:-1]), x1.shape[-1]
    F = h.shape[-1]
    dug, dxn = _call("MpxGluBwdFused", (_sds((M, 2 * F), BF16), _sds((M, H), BF16)), xn, dy, wt_ug64, wt_d)
    d_down = _wgrad_raw(h, dy)
    d_up, d_gate = _wgrad_raw(xn, dug[:, :F]), _wgrad_raw(xn, dug[:, F:])
    dx1, dscale = _call("MpxRmsNormBwd", (_sds(x1.shape, BF16), _sds((H,), F32)), dxn.reshape(x1.shape), x1, norm_scale, dy,
                        jnp.zeros((H,), F32), aliases={4: 1}, eps=np.float32(1e-6))
    return dx1, dscale, d_up, d_gate, d_down


glu_block.defvjp(_glu_fwd, _glu_bwd)


def ffn_half_decode(attn_out, x, norm_scale, wt_ug64, wt_d, wt_o, eps: float = 1e-6):
    """Rollout: out projection + residual + ffn_norm + GLU + residual in one kernel (network.py:163-170)."""
    out, _ = _call("MpxGluFused", (_sds(x.shape, BF16), _sds((0,), BF16)), attn_out, wt_ug64, wt_d, x, norm_scale, wt_o,
                   eps=np.float32(eps))
    return out


def ffn_half_decode_with_policy_head(attn_out, x, norm_scale, wt_ug64, wt_d, wt_o, out_norm_scale, table, action_mask, gumbel,
                                     eps: float = 1e-6):
    """Last layer of a rollout step: ffn_half_decode + policy_head_sample in one launch (bit-identical results).
    -> (block output, action, log_prob, logits)."""
    B, A = x.shape[0], table.shape[0]
    mask = jnp.zeros((0,), jnp.bool_) if action_mask is None else action_mask
    out, action, log_prob, logits, _ = _call(
        "MpxGluFusedHead", (_sds(x.shape, BF16), _sds((B,), I32), _sds((B,), F32), _sds((B, A), F32), _sds((0,), BF16)),
        attn_out, wt_ug64, wt_d, x, norm_scale, wt_o, out_norm_scale, table, mask, gumbel, jnp.zeros((0,), jnp.uint64),
        eps=np.float32(eps), head_eps=np.float32(eps), seed=np.int64(0), stream_id=np.int64(0))
    return out, action, log_prob, logits


# ------------------------------------------------------------------------------------------------ rollout heads
def obs_fused_prep(sizes, w1, b1, wt2, b2, wt3, b3):
    return _call("MpxObsFusedPrep", _sds((_CORE.mpx_obs_fused_image_bytes(),), U8), w1, b1, wt2, b2, wt3, b3,
                 sizes=np.asarray(sizes, np.int32))


def obs_embed_fused(obs, image, reward, last_action, task_ids, w_reward, b_reward, action_table, task_table, sizes, H: int = 128):
    B = obs.shape[0]
    none_i = jnp.zeros((0,), I32)
    return _call("MpxObsEmbedFused", _sds((B, H), BF16), obs, image, reward, last_action,
                 none_i if task_ids is None else task_ids, w_reward, b_reward, action_table,
                 _EMPTY_F32() if task_table is None else task_table, sizes=np.asarray(sizes, np.int32))


# ------------------------------------------------------------------------------------------------ training encoder
def _obs_tiles(M: int) -> int:
    return (M + 127) // 128


def obs_train_fwd(obs, image, reward, last_action, task_ids, w_reward, b_reward, action_table, task_table, sizes, H: int = 128):
    """GridCnnObsEncoder + embedding sum under nnx.grad (observation.py:84-96, network.py:303-309) in one kernel.
    Returns x (M,H) and the residuals (z1t, a1t, z2t, a2) the backward consumes."""
    M = math.prod(obs.shape[:-3])
    nt = _obs_tiles(M)
    none_i = jnp.zeros((0,), I32)
    return _call("MpxObsTrainFwd", (_sds((M, H), BF16), _sds((nt * 25 * 2048,), BF16), _sds((nt * 25 * 2048,), BF16),
                                    _sds((nt * 9 * 4096,), BF16), _sds((M, 288), BF16)),
                 obs, image, reward, last_action, none_i if task_ids is None else task_ids, w_reward, b_reward, action_table,
                 _EMPTY_F32() if task_table is None else task_table, sizes=np.asarray(sizes, np.int32))


def obs_train_bwd(dx0, z1t, z2t, image, db1, db2):
    """Data gradients of the encoder given dx0 = d loss / d (encoder output): (dz2t, dz1, db1 + d b1, db2 + d b2)."""
    M = dx0.shape[0]
    dz2t, dz1, db1, db2, _ = _call("MpxObsTrainBwd", (_sds(z2t.shape, BF16), _sds((M * 25, 16), BF16), _sds((16,), F32),
                                                      _sds((32,), F32), _sds((_CORE.mpx_obs_train_bwd_workspace_bytes(),), U8)),
                                   dx0, z1t, z2t, image, db1, db2, aliases={4: 2, 5: 3})
    return dz2t, dz1, db1, db2


def obs_conv2_wgrad(a1t, dz2t, dw, M: int):
    """dw (3,3,16,32) f32 += the tap-wise conv2 weight gradient (no im2col)."""
    dw, _ = _call("MpxObsConv2Wgrad", (_sds(dw.shape, F32), _sds((_CORE.mpx_obs_train_bwd_workspace_bytes(),), U8)),
                  a1t, dz2t, dw, aliases={2: 0}, M=np.int64(M))
    return dw


def policy_head_sample(x, norm_scale, table, action_mask, gumbel, eps: float = 1e-6):
    """out norm + tied logits + mask + Gumbel arg-max + log-prob (network.py:321-333, train.py:97-98).  `gumbel` is the
    noise jax.random.gumbel(key, (B,A)) draws - the same noise distrax would use, so the sampled index is bit-identical
    to argmax(logits + gumbel) of the reference for equal logits."""
    B, A = x.shape[0], table.shape[0]
    mask = jnp.zeros((0,), jnp.bool_) if action_mask is None else action_mask
    action, log_prob, logits = _call("MpxPolicyHeadSample", (_sds((B,), I32), _sds((B,), F32), _sds((B, A), F32)),
                                     x, norm_scale, table, mask, gumbel, jnp.zeros((0,), jnp.uint64),
                                     eps=np.float32(eps), seed=np.int64(0), stream_id=np.int64(0))
    return action, log_prob, logits


def value_head_fused(x, norm_scale, wt_vm, b_vm, wt_vh_hilo, b_vh, centers, eps: float = 1e-6):
    """value MLP + fp32 HL-Gauss head + expectation of a rollout step (network.py:321,335-341, values.py:40-45)."""
    B, NL = x.shape[0], centers.shape[0]
    return _call("MpxValueHeadFused", (_sds((B,), F32), _sds((B, NL), F32)), x, norm_scale, wt_vm, b_vm, wt_vh_hilo, b_vh,
                 centers, eps=np.float32(eps))


def value_bwd_fused(xf, dvl, wt_vm, b_vm, w_vh_pad, db_vm):
    """Backward of value_mlp + value_head in one kernel (network.py:335-341, values.py:34,40-41): returns
    (pv, dz, dxf, db_vm + d b_vm); dWvh = pv^T dvl and dWvm = xf^T dz are linear_wgrad calls on pv / dz."""
    M, H = xf.shape
    Vh = wt_vm.shape[0]
    pv, dz, dxf, db, _ = _call("MpxValueBwdFused", (_sds((M, Vh), BF16), _sds((M, Vh), BF16), _sds((M, H), BF16),
                                                    _sds((Vh,), F32), _sds((_CORE.mpx_value_bwd_fused_workspace_bytes(Vh),), U8)),
                               xf, dvl, wt_vm, b_vm, w_vh_pad, db_vm, aliases={5: 3})
    return pv, dz, dxf, db
