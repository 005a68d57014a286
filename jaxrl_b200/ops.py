"""Torch-tensor front end of the C ABI: one python function per entry point of include/mapox_b200.h.

torch is used for device memory and the current CUDA stream only; every arithmetic step runs inside
libmapox_b200.so.  All tensors must live on a CUDA device (there is no CPU path).
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check, lib, ptr, stream

BF16 = torch.bfloat16
F32 = torch.float32


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.MpxError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.MpxError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.MpxError(f"{name}: expected a contiguous tensor")
    return t


# ------------------------------------------------------------------------------------------------ GAE
def gae(rewards, values, next_terminated, discount: float, gae_lambda: float):
    """-> (advantages, targets, stats[2] f64).  rollout.py:128-157 (before normalisation)."""
    B, T = rewards.shape
    _req(rewards, F32, "rewards"); _req(values, F32, "values")
    if next_terminated.dtype == torch.bool:
        next_terminated = next_terminated.view(torch.uint8)
    _req(next_terminated, torch.uint8, "next_terminated")
    assert values.shape == (B, T + 1)
    adv = torch.empty_like(rewards)
    tgt = torch.empty_like(rewards)
    stats = torch.empty(2, dtype=torch.float64, device=rewards.device)
    ws = torch.empty(lib.mpx_gae_workspace_bytes(B, T), dtype=torch.uint8, device=rewards.device)
    check(lib.mpx_gae(ptr(rewards), ptr(values), ptr(next_terminated), B, T, discount, gae_lambda,
                      ptr(adv), ptr(tgt), ptr(stats), ptr(ws), stream()), "mpx_gae")
    return adv, tgt, stats


def adv_normalize(adv, stats, count: Optional[float] = None):
    """In place (adv - mean) / (std + 1e-8), rollout.py:159-162."""
    _req(adv, F32, "adv"); _req(stats, torch.float64, "stats")
    n = adv.numel()
    check(lib.mpx_adv_normalize(ptr(adv), n, ptr(stats), float(n if count is None else count), stream()),
          "mpx_adv_normalize")
    return adv


# ------------------------------------------------------------------------------------------------ HL-Gauss
def hlgauss_value(logits, centers):
    _req(logits, F32, "logits"); _req(centers, F32, "centers")
    NL = logits.shape[-1]
    n = logits.numel() // NL
    out = torch.empty(logits.shape[:-1], dtype=F32, device=logits.device)
    check(lib.mpx_hlgauss_value(ptr(logits), n, NL, ptr(centers), ptr(out), stream()), "mpx_hlgauss_value")
    return out


def hlgauss_loss(logits, targets, supports, vmin: float, vmax: float, sigma: float,
                 grad_scale: float = 1.0, want_grad: bool = True):
    """-> (loss[1] f32, dlogits or None). values.py:47-63."""
    _req(logits, F32, "logits"); _req(targets, F32, "targets"); _req(supports, F32, "supports")
    NL = logits.shape[-1]
    n = logits.numel() // NL
    assert targets.numel() == n and supports.numel() == NL + 1
    loss = torch.empty(1, dtype=F32, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    ws = torch.empty(lib.mpx_hlgauss_loss_workspace_bytes(n), dtype=torch.uint8, device=logits.device)
    check(lib.mpx_hlgauss_loss(ptr(logits), ptr(targets), n, NL, ptr(supports), vmin, vmax, sigma, grad_scale,
                               ptr(loss), ptr(dl), ptr(ws), stream()), "mpx_hlgauss_loss")
    return loss, dl


# ------------------------------------------------------------------------------------------------ dense
def linear(x, wt, bias=None, residual=None, act: int = 0, out: Optional[torch.Tensor] = None,
           out_f32: Optional[torch.Tensor] = None, want_bf16: bool = True):
    """y = epi(x @ wt^T): x (M,K) bf16, wt (N,K) bf16 (transposed flax kernel)."""
    _req(x, BF16, "x"); _req(wt, BF16, "wt")
    M, K = x.shape
    N = wt.shape[0]
    assert wt.shape[1] == K
    if want_bf16 and out is None:
        out = torch.empty(M, N, dtype=BF16, device=x.device)
    a = _lib.LinearArgs(ptr(x), K, ptr(wt), K, ptr(bias), ptr(residual), N, ptr(out), N,
                        ptr(out_f32), N, M, N, K, act)
    check(lib.mpx_linear(C.byref(a), stream()), "mpx_linear")
    return out if want_bf16 else out_f32


def linear_wgrad(x, dy, dw):
    """dw (K,N) f32 += x^T @ dy."""
    _req(x, BF16, "x"); _req(dy, BF16, "dy"); _req(dw, F32, "dw")
    M, K = x.shape
    N = dy.shape[1]
    assert dy.shape[0] == M and dw.shape == (K, N)
    a = _lib.WgradArgs(ptr(x), K, ptr(dy), N, ptr(dw), N, M, N, K)
    check(lib.mpx_linear_wgrad(C.byref(a), stream()), "mpx_linear_wgrad")
    return dw


def qkv_rope(x, wt_qkv, pos, timescale, Nq: int, Nkv: int, D: int, q=None, k=None, v=None,
             k_row_stride: Optional[int] = None, v_row_stride: Optional[int] = None, cache_S: int = 0):
    """q/k/v projection + RoPE (+ ring-slot write when cache_S > 0).  Returns (q, k, v)."""
    _req(x, BF16, "x"); _req(wt_qkv, BF16, "wt_qkv"); _req(pos, torch.int32, "pos"); _req(timescale, F32, "timescale")
    M, H = x.shape
    if q is None:
        q = torch.empty(M, Nq * D, dtype=BF16, device=x.device)
    if k is None:
        k = torch.empty(M, Nkv * D, dtype=BF16, device=x.device)
        v = torch.empty(M, Nkv * D, dtype=BF16, device=x.device)
        k_row_stride = v_row_stride = Nkv * D
    a = _lib.QkvRopeArgs(ptr(x), H, ptr(wt_qkv), ptr(pos), pos.numel(), ptr(timescale), ptr(q), ptr(k),
                         k_row_stride, ptr(v), v_row_stride, cache_S, M, H, Nq, Nkv, D)
    check(lib.mpx_qkv_rope(C.byref(a), stream()), "mpx_qkv_rope")
    return q, k, v


def glu_up(x, wt_ug, F: int, save: bool = False):
    """h = bf16(bf16(gelu(x@Wup)) * bf16(x@Wgate)); optionally also returns the bf16 pre-activations."""
    _req(x, BF16, "x"); _req(wt_ug, BF16, "wt_ug")
    M, K = x.shape
    h = torch.empty(M, F, dtype=BF16, device=x.device)
    u = torch.empty(M, F, dtype=BF16, device=x.device) if save else None
    g = torch.empty(M, F, dtype=BF16, device=x.device) if save else None
    a = _lib.GluUpArgs(ptr(x), K, ptr(wt_ug), ptr(h), ptr(u), ptr(g), M, F, K)
    check(lib.mpx_glu_up(C.byref(a), stream()), "mpx_glu_up")
    return (h, u, g) if save else h


def pack_glu_weight(w_up: torch.Tensor, w_gate: torch.Tensor) -> torch.Tensor:
    """(K,F) flax kernels -> (2F,K) bf16, 256-row tiles [128 up rows | 128 gate rows] (see mpx_glu_up)."""
    K, F = w_up.shape
    assert F % 128 == 0
    ut = w_up.t().reshape(F // 128, 128, K)
    gt = w_gate.t().reshape(F // 128, 128, K)
    return torch.cat([ut, gt], dim=1).reshape(2 * F, K).to(BF16).contiguous()


# ------------------------------------------------------------------------------------------------ attention
def attn_decode(q, kcache, vcache, pos, Nq: int, Nkv: int, D: int, out=None):
    _req(q, BF16, "q"); _req(kcache, BF16, "kcache"); _req(vcache, BF16, "vcache"); _req(pos, torch.int32, "pos")
    B, S = kcache.shape[0], kcache.shape[1]
    if out is None:
        out = torch.empty(B, Nq * D, dtype=BF16, device=q.device)
    a = _lib.AttnDecodeArgs(ptr(q), ptr(kcache), ptr(vcache), ptr(pos), ptr(out), B, S, Nq, Nkv, D)
    check(lib.mpx_attn_decode(C.byref(a), stream()), "mpx_attn_decode")
    return out
