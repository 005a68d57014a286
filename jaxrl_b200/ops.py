"""Torch-tensor front end of the C ABI: one python function per entry point of include/mapox_b200.h.

torch is used for device memory and the current CUDA stream only; every arithmetic step runs inside
libmapox_b200.so.  All tensors must live on a CUDA device (there is no CPU path).
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, lib, ptr, stream

BF16 = torch.bfloat16
F32 = torch.float32
I32 = torch.int32

# kernels launched per C-ABI call where it is not one (bench.py counts `gpu_launches` from the calls of one eager update);
# muon: prepare + 3 x 5 Newton-Schulz GEMMs + finish (phase 1), AdamW leaves + clip/apply (phase 2)
LAUNCHES_PER_CALL = {"mpx_gae": 2, "mpx_hlgauss_loss": 2, "mpx_ppo_policy_loss": 2, "mpx_adamw_step": 2,
                     "mpx_attn_train_bwd": 3, "mpx_muon_step:1": 17, "mpx_muon_step:2": 2, "mpx_muon_step:3": 19,
                     "mpx_muon_apply_sharded": 3,
                     "mpx_linear_wgrad": 2}


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.MpxError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.MpxError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.MpxError(f"{name}: expected a contiguous tensor")
    return t


def _u8(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.view(torch.uint8) if t.dtype == torch.bool else t


# ------------------------------------------------------------------------------------------------ GAE
def gae(rewards, values, next_terminated, discount: float, gae_lambda: float, adv=None, tgt=None, stats=None,
        ws=None):
    """-> (advantages, targets, stats[2] f64).  rollout.py:128-157 (before normalisation)."""
    B, T = rewards.shape
    _req(rewards, F32, "rewards"); _req(values, F32, "values")
    next_terminated = _req(_u8(next_terminated), torch.uint8, "next_terminated")
    assert values.shape == (B, T + 1)
    adv = torch.empty_like(rewards) if adv is None else adv
    tgt = torch.empty_like(rewards) if tgt is None else tgt
    stats = torch.empty(2, dtype=torch.float64, device=rewards.device) if stats is None else stats
    if ws is None:
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
def hlgauss_value(logits, centers, out=None):
    _req(logits, F32, "logits"); _req(centers, F32, "centers")
    NL = logits.shape[-1]
    n = logits.numel() // NL
    out = torch.empty(logits.shape[:-1], dtype=F32, device=logits.device) if out is None else out
    check(lib.mpx_hlgauss_value(ptr(logits), n, NL, ptr(centers), ptr(out), stream()), "mpx_hlgauss_value")
    return out


def hlgauss_loss(logits, targets, supports, vmin: float, vmax: float, sigma: float, grad_scale: float = 1.0,
                 want_grad: bool = True, loss=None, dlogits=None, dlogits_bf16=None, ws=None):
    """-> (loss[1] f32, dlogits or None). values.py:47-63.  dlogits_bf16 (n, ld) optional padded bf16 gradient."""
    _req(logits, F32, "logits"); _req(targets, F32, "targets"); _req(supports, F32, "supports")
    NL = logits.shape[-1]
    n = logits.numel() // NL
    assert targets.numel() == n and supports.numel() == NL + 1
    loss = torch.empty(1, dtype=F32, device=logits.device) if loss is None else loss
    if dlogits is None and want_grad and dlogits_bf16 is None:
        dlogits = torch.empty_like(logits)
    if ws is None:
        ws = torch.empty(lib.mpx_hlgauss_loss_workspace_bytes(n), dtype=torch.uint8, device=logits.device)
    ld = 0 if dlogits_bf16 is None else dlogits_bf16.shape[-1]
    check(lib.mpx_hlgauss_loss(ptr(logits), ptr(targets), n, NL, ptr(supports), vmin, vmax, sigma, grad_scale,
                               ptr(loss), ptr(dlogits), ptr(dlogits_bf16), ld, ptr(ws), stream()), "mpx_hlgauss_loss")
    return loss, dlogits


# ------------------------------------------------------------------------------------------------ dense
def linear(x, wt, bias=None, residual=None, act: int = 0, out: Optional[torch.Tensor] = None,
           out_f32: Optional[torch.Tensor] = None, want_bf16: bool = True, y_pre=None, M: Optional[int] = None):
    """y = epi(x @ wt^T): x (M,K) bf16, wt (N,K) bf16 (transposed flax kernel)."""
    _req(x, BF16, "x"); _req(wt, BF16, "wt")
    K = x.shape[-1]
    M = x.numel() // K if M is None else M
    N = wt.shape[0]
    assert wt.shape[1] == K
    if want_bf16 and out is None:
        out = torch.empty(M, N, dtype=BF16, device=x.device)
    a = _lib.LinearArgs(ptr(x), K, ptr(wt), K, ptr(bias), ptr(residual), N, ptr(out), N,
                        ptr(out_f32), N, M, N, K, act, ptr(y_pre))
    check(lib.mpx_linear(C.byref(a), stream()), "mpx_linear")
    return out if want_bf16 else out_f32


_WGRAD_WS = {}


def wgrad_workspace(device, nbytes: int = 96 << 20) -> torch.Tensor:
    """Shared scratch for the deterministic split-K reduction of mpx_linear_wgrad (one per device)."""
    key = (torch.device(device).index or 0)
    ws = _WGRAD_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WGRAD_WS[key] = ws
    return ws


def linear_wgrad(x, dy, dw, M: Optional[int] = None, K: Optional[int] = None, N: Optional[int] = None,
                 ldx: Optional[int] = None, ldy: Optional[int] = None, lddw: Optional[int] = None, ws=None):
    """dw (K,N) f32 += x^T @ dy.  Column sub-blocks are addressed with explicit pitches (ldx/ldy/lddw)."""
    _req(dw, F32, "dw")
    assert x.dtype == BF16 and dy.dtype == BF16 and x.is_cuda and dy.is_cuda
    ldx = x.shape[-1] if ldx is None else ldx
    ldy = dy.shape[-1] if ldy is None else ldy
    K = ldx if K is None else K
    N = ldy if N is None else N
    M = x.numel() // x.shape[-1] if M is None else M
    lddw = dw.shape[-1] if lddw is None else lddw
    if ws is None:
        ws = wgrad_workspace(dw.device)
    a = _lib.WgradArgs(ptr(x), ldx, ptr(dy), ldy, ptr(dw), lddw, M, N, K, ptr(ws), ws.numel())
    check(lib.mpx_linear_wgrad(C.byref(a), stream()), "mpx_linear_wgrad")
    return dw


def linear_f32w(x, wt_hilo, bias, out, N: int):
    """y (M,N) f32 = f32(x) @ W + bias with W carried as a bf16 hi/lo pair (HL-Gauss head)."""
    _req(x, BF16, "x"); _req(wt_hilo, BF16, "wt_hilo"); _req(out, F32, "out")
    K = x.shape[-1]
    M = x.numel() // K
    check(lib.mpx_linear_f32w(ptr(x), K, ptr(wt_hilo), ptr(bias), ptr(out), out.shape[-1], M, N, K, stream()),
          "mpx_linear_f32w")
    return out


def qkv_rope(x, wt_qkv, pos, timescale, Nq: int, Nkv: int, D: int, q=None, k=None, v=None,
             q_row_stride: Optional[int] = None, k_row_stride: Optional[int] = None,
             v_row_stride: Optional[int] = None, cache_S: int = 0, pos_len: Optional[int] = None, norm_scale=None,
             eps: float = 1e-6):
    """q/k/v projection + RoPE (+ ring-slot write when cache_S > 0), optionally preceded by RMSNorm(x; norm_scale).
    Returns (q, k, v)."""
    _req(x, BF16, "x"); _req(wt_qkv, BF16, "wt_qkv"); _req(timescale, F32, "timescale")
    assert pos.dtype == I32 and pos.is_cuda
    H = x.shape[-1]
    M = x.numel() // H
    if q is None:
        q = torch.empty(M, Nq * D, dtype=BF16, device=x.device)
    if k is None:
        k = torch.empty(M, Nkv * D, dtype=BF16, device=x.device)
        v = torch.empty(M, Nkv * D, dtype=BF16, device=x.device)
    q_row_stride = Nq * D if q_row_stride is None else q_row_stride
    k_row_stride = Nkv * D if k_row_stride is None else k_row_stride
    v_row_stride = Nkv * D if v_row_stride is None else v_row_stride
    a = _lib.QkvRopeArgs(ptr(x), H, ptr(wt_qkv), ptr(pos), pos.numel() if pos_len is None else pos_len,
                         ptr(timescale), ptr(q), q_row_stride, ptr(k), k_row_stride, ptr(v), v_row_stride,
                         cache_S, M, H, Nq, Nkv, D, ptr(norm_scale), eps)
    check(lib.mpx_qkv_rope(C.byref(a), stream()), "mpx_qkv_rope")
    return q, k, v


def glu_up(x, wt_ug, F: int, save: bool = False, h=None, u=None, g=None):
    """h = bf16(bf16(gelu(x@Wup)) * bf16(x@Wgate)); optionally also returns the bf16 pre-activations."""
    _req(x, BF16, "x"); _req(wt_ug, BF16, "wt_ug")
    K = x.shape[-1]
    M = x.numel() // K
    h = torch.empty(M, F, dtype=BF16, device=x.device) if h is None else h
    if save and u is None:
        u = torch.empty(M, F, dtype=BF16, device=x.device)
        g = torch.empty(M, F, dtype=BF16, device=x.device)
    a = _lib.GluUpArgs(ptr(x), K, ptr(wt_ug), ptr(h), ptr(u), ptr(g), M, F, K)
    check(lib.mpx_glu_up(C.byref(a), stream()), "mpx_glu_up")
    return (h, u, g) if save else h


def glu_fused(x, wt_ug64, wt_d, residual, out, F: int, norm_scale=None, eps: float = 1e-6, wt_o=None, h_out=None):
    """out = residual + down(gelu(up(xn)) * gate(xn)) in one kernel (H = 128); xn = RMSNorm(x) when norm_scale is given.
    With wt_o (Wo^T): x is the attention output, x1 = residual + x @ Wo, out = x1 + GLU(RMSNorm(x1)).
    h_out (M,F), if given, also receives the hidden activation h (kept by the training forward for dWdown)."""
    _req(x, BF16, "x"); _req(wt_ug64, BF16, "wt_ug64"); _req(wt_d, BF16, "wt_d"); _req(residual, BF16, "residual")
    H = x.shape[-1]
    M = x.numel() // H
    check(lib.mpx_glu_fused(ptr(x), ptr(wt_ug64), ptr(wt_d), ptr(residual), ptr(out), M, H, F, ptr(norm_scale), eps,
                            ptr(wt_o), ptr(h_out), stream()), "mpx_glu_fused")
    return out


def glu_fused_head(x, wt_ug64, wt_d, residual, out, F: int, norm_scale, wt_o, head_norm_scale, table, action, log_prob,
                   seed: int = 0, stream_id: int = 0, stream_offset=None, mask=None, gumbel_in=None, logits_out=None,
                   xf_out=None, eps: float = 1e-6, head_eps: float = 1e-6):
    """The last layer's feed-forward half with the policy head folded in: glu_fused(..., wt_o=wt_o) followed by
    policy_head_sample(out, ...) as ONE launch (bit-identical results; the library issues the two kernels itself for shapes
    the fused epilogue does not cover)."""
    _req(x, BF16, "x"); _req(wt_ug64, BF16, "wt_ug64"); _req(wt_d, BF16, "wt_d"); _req(residual, BF16, "residual")
    _req(table, F32, "table"); _req(head_norm_scale, F32, "head_norm_scale")
    H = x.shape[-1]
    M = x.numel() // H
    h = _lib.PolicyHeadArgs(ptr(head_norm_scale), head_eps, ptr(table), table.shape[0], ptr(_u8(mask)), ptr(gumbel_in), seed,
                            stream_id, ptr(stream_offset), ptr(logits_out), ptr(xf_out), ptr(action), ptr(log_prob))
    check(lib.mpx_glu_fused_head(ptr(x), ptr(wt_ug64), ptr(wt_d), ptr(residual), ptr(out), M, H, F, ptr(norm_scale), eps,
                                 ptr(wt_o), C.byref(h), stream()), "mpx_glu_fused_head")
    return out


def glu_bwd_fused(xn, dy, wt_ug64, wt_d, F: int, dug=None, dxn=None):
    """Backward of the GLU block for a forward that kept only xn (and h): recomputes u, g, dh on chip; returns
    (dug, dxn) with dug = [du | dg] (M, 2F) - the operand of the up/gate weight gradients - and dxn (M, H).  H = 128."""
    _req(xn, BF16, "xn"); _req(dy, BF16, "dy"); _req(wt_ug64, BF16, "wt_ug64"); _req(wt_d, BF16, "wt_d")
    H = xn.shape[-1]
    M = xn.numel() // H
    dug = torch.empty(M, 2 * F, dtype=BF16, device=xn.device) if dug is None else dug
    dxn = torch.empty(M, H, dtype=BF16, device=xn.device) if dxn is None else dxn
    check(lib.mpx_glu_bwd_fused(ptr(xn), ptr(dy), ptr(wt_ug64), ptr(wt_d), ptr(dug), ptr(dxn), M, H, F, stream()),
          "mpx_glu_bwd_fused")
    return dug, dxn


def value_head_fused(x, wt_vm, b_vm, wt_vh_hilo, b_vh, centers, value, vlogits=None, norm_scale=None,
                     eps: float = 1e-6):
    """value = HL-Gauss expectation of the value head on gelu(value_mlp(x)) in one kernel (H = 128)."""
    _req(x, BF16, "x"); _req(wt_vm, BF16, "wt_vm"); _req(b_vm, BF16, "b_vm"); _req(wt_vh_hilo, BF16, "wt_vh_hilo")
    _req(b_vh, F32, "b_vh"); _req(centers, F32, "centers"); _req(value, F32, "value")
    H = x.shape[-1]
    M = x.numel() // H
    Vh = wt_vm.shape[0]
    NL = centers.numel()
    check(lib.mpx_value_head_fused(ptr(x), ptr(norm_scale), eps, ptr(wt_vm), ptr(b_vm), ptr(wt_vh_hilo), ptr(b_vh),
                                   ptr(centers), ptr(value), ptr(vlogits), M, H, Vh, NL, stream()),
          "mpx_value_head_fused")
    return value


def value_bwd_fused(xf, dvl, wt_vm, b_vm, w_vh_pad, pv, dz, dxf, db_vm, ws=None):
    """Backward of the value MLP + head in one kernel: recomputes the pre-activation, writes pv / dz (weight-gradient
    operands), dxf, and accumulates the value-MLP bias gradient."""
    for t, n in ((xf, "xf"), (dvl, "dvl"), (wt_vm, "wt_vm"), (b_vm, "b_vm"), (w_vh_pad, "w_vh_pad"), (pv, "pv"), (dz, "dz"),
                 (dxf, "dxf")):
        _req(t, BF16, n)
    _req(db_vm, F32, "db_vm")
    H = xf.shape[-1]
    M = xf.numel() // H
    Vh = wt_vm.shape[0]
    if dvl.shape[-1] != 64 or w_vh_pad.shape[-1] != 64:
        raise _lib.MpxError("value_bwd_fused: dvl / w_vh_pad must have 64 (padded) logit columns")
    need = lib.mpx_value_bwd_fused_workspace_bytes(Vh)
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device=xf.device)
    check(lib.mpx_value_bwd_fused(ptr(xf), ptr(dvl), ptr(wt_vm), ptr(b_vm), ptr(w_vh_pad), ptr(pv), ptr(dz), ptr(dxf),
                                  ptr(db_vm), M, H, Vh, ptr(ws), ws.numel() * ws.element_size(), stream()),
          "mpx_value_bwd_fused")
    return dxf


def pack_glu_weight64(w_up: torch.Tensor, w_gate: torch.Tensor) -> torch.Tensor:
    """(K,F) flax kernels -> (2F,K) bf16, 128-row tiles [64 up rows | 64 gate rows] (see mpx_glu_fused)."""
    K, F = w_up.shape
    assert F % 64 == 0
    ut = w_up.t().reshape(F // 64, 64, K)
    gt = w_gate.t().reshape(F // 64, 64, K)
    return torch.cat([ut, gt], dim=1).reshape(2 * F, K).to(BF16).contiguous()


def pack_glu_weight(w_up: torch.Tensor, w_gate: torch.Tensor) -> torch.Tensor:
    """(K,F) flax kernels -> (2F,K) bf16, 256-row tiles [128 up rows | 128 gate rows] (see mpx_glu_up)."""
    K, F = w_up.shape
    assert F % 128 == 0
    ut = w_up.t().reshape(F // 128, 128, K)
    gt = w_gate.t().reshape(F // 128, 128, K)
    return torch.cat([ut, gt], dim=1).reshape(2 * F, K).to(BF16).contiguous()


# ------------------------------------------------------------------------------------------------ row-wise
def rmsnorm_fwd(x, scale, out=None, eps: float = 1e-6):
    _req(x, BF16, "x"); _req(scale, F32, "scale")
    H = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    check(lib.mpx_rmsnorm_fwd(ptr(x), ptr(scale), eps, ptr(out), x.numel() // H, H, stream()), "mpx_rmsnorm_fwd")
    return out


def rmsnorm_bwd(dy, x, scale, dscale=None, dres=None, out=None, eps: float = 1e-6):
    _req(dy, BF16, "dy"); _req(x, BF16, "x"); _req(scale, F32, "scale")
    H = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    check(lib.mpx_rmsnorm_bwd(ptr(dy), ptr(x), ptr(scale), eps, ptr(dres), ptr(out), ptr(dscale), x.numel() // H, H,
                              stream()), "mpx_rmsnorm_bwd")
    return out


def glu_bwd(dh, u, g, out=None):
    F = dh.shape[-1]
    M = dh.numel() // F
    out = torch.empty(M, 2 * F, dtype=BF16, device=dh.device) if out is None else out
    check(lib.mpx_glu_bwd(ptr(dh), ptr(u), ptr(g), ptr(out), M, F, stream()), "mpx_glu_bwd")
    return out


def gelu_bwd(dy, z, out=None):
    out = torch.empty_like(dy) if out is None else out
    check(lib.mpx_gelu_bwd(ptr(dy), ptr(z), ptr(out), dy.numel(), stream()), "mpx_gelu_bwd")
    return out


def rope_(x, ld: int, pos, timescale, M: int, nheads: int, head_dim: int = 32, inverse: bool = False,
          pos_len: Optional[int] = None):
    check(lib.mpx_rope(ptr(x), ld, ptr(pos), pos.numel() if pos_len is None else pos_len, ptr(timescale), M, nheads,
                       head_dim, int(inverse), stream()), "mpx_rope")
    return x


def colsum(x, out, M: Optional[int] = None, N: Optional[int] = None, ld: Optional[int] = None):
    ld = x.shape[-1] if ld is None else ld
    N = ld if N is None else N
    M = x.numel() // x.shape[-1] if M is None else M
    check(lib.mpx_colsum(ptr(x), ld, M, N, ptr(out), stream()), "mpx_colsum")
    return out


def add_bf16(a, b, out=None):
    out = torch.empty_like(a) if out is None else out
    check(lib.mpx_add_bf16(ptr(a), ptr(b), ptr(out), a.numel(), stream()), "mpx_add_bf16")
    return out


# ------------------------------------------------------------------------------------------------ attention
def attn_decode(q, kcache, vcache, pos, Nq: int, Nkv: int, D: int, out=None):
    _req(q, BF16, "q"); _req(kcache, BF16, "kcache"); _req(vcache, BF16, "vcache")
    assert pos.dtype == I32 and pos.is_cuda
    B, S = kcache.shape[0], kcache.shape[1]
    if out is None:
        out = torch.empty(B, Nq * D, dtype=BF16, device=q.device)
    a = _lib.AttnDecodeArgs(ptr(q), ptr(kcache), ptr(vcache), ptr(pos), ptr(out), B, S, Nq, Nkv, D)
    check(lib.mpx_attn_decode(C.byref(a), stream()), "mpx_attn_decode")
    return out


def attn_decode_fused(x, norm_scale, wt_qkv, pos, timescale, kcache, vcache, Nq: int, Nkv: int, D: int, out=None,
                      eps: float = 1e-6, stable_history: bool = False):
    """history_norm + q/k/v projection + RoPE + ring write + decode attention in one kernel (H in {128, 256}, Nq=4, Nkv=1, D=32).
    kcache/vcache (B,S,Nkv,D) are updated in place at slot pos[0] % S; returns out (B, Nq*D).
    stable_history: the caller guarantees that the other ring slots were written by kernels that completed before this
    launch could start (flags bit 0 of the C ABI) - lets the cache stream start under the previous kernel's tail."""
    _req(x, BF16, "x"); _req(wt_qkv, BF16, "wt_qkv"); _req(kcache, BF16, "kcache"); _req(vcache, BF16, "vcache")
    _req(norm_scale, F32, "norm_scale"); _req(timescale, F32, "timescale")
    assert pos.dtype == I32 and pos.is_cuda
    B, S = kcache.shape[0], kcache.shape[1]
    H = x.shape[-1]
    if out is None:
        out = torch.empty(B, Nq * D, dtype=BF16, device=x.device)
    a = _lib.AttnDecodeFusedArgs(ptr(x), ptr(norm_scale), eps, ptr(wt_qkv), ptr(pos), ptr(timescale), ptr(kcache),
                                 ptr(vcache), ptr(out), B, S, H, Nq, Nkv, D, 1 if stable_history else 0)
    check(lib.mpx_attn_decode_fused(C.byref(a), stream()), "mpx_attn_decode_fused")
    return out


def attn_train_fwd(q, k, v, B: int, T: int, Nq: int, Nkv: int, D: int, ldq: int, ldk: int, ldv: int, o=None,
                   lse=None):
    """q/k/v may be column slices (data_ptr offset views) of one (B*T, ld) buffer."""
    M = B * T
    o = torch.empty(M, Nq * D, dtype=BF16, device=q.device) if o is None else o
    lse = torch.empty(M, Nq, dtype=F32, device=q.device) if lse is None else lse
    a = _lib.AttnTrainArgs(ptr(q), ldq, ptr(k), ldk, ptr(v), ldv, ptr(o), Nq * D, ptr(lse), B, T, Nq, Nkv, D)
    check(lib.mpx_attn_train_fwd(C.byref(a), stream()), "mpx_attn_train_fwd")
    return o, lse


def attn_train_bwd(q, k, v, o, dout, lse, pos, timescale, B: int, T: int, Nq: int, Nkv: int, D: int, ldq: int,
                   ldk: int, ldv: int, dq, dk, dv, lddq: int, lddk: int, lddv: int, delta=None, dq_acc=None,
                   pos_len: Optional[int] = None):
    M = B * T
    delta = torch.empty(M, Nq, dtype=F32, device=q.device) if delta is None else delta
    dq_acc = torch.empty(M, Nq * D, dtype=F32, device=q.device) if dq_acc is None else dq_acc
    a = _lib.AttnTrainBwdArgs(ptr(q), ldq, ptr(k), ldk, ptr(v), ldv, ptr(o), Nq * D, ptr(dout), Nq * D, ptr(lse),
                              ptr(delta), ptr(dq_acc), ptr(dq), lddq, ptr(dk), lddk, ptr(dv), lddv, ptr(pos),
                              pos.numel() if pos_len is None else pos_len, ptr(timescale), B, T, Nq, Nkv, D)
    check(lib.mpx_attn_train_bwd(C.byref(a), stream()), "mpx_attn_train_bwd")
    return dq, dk, dv


# ------------------------------------------------------------------------------------------------ embeds / heads
def embed_combine(obs_emb, reward, last_action, task_ids, w_r, b_r, a_table, t_table, out=None):
    H = obs_emb.shape[-1]
    M = obs_emb.numel() // H
    out = torch.empty_like(obs_emb) if out is None else out
    check(lib.mpx_embed_combine(ptr(obs_emb), ptr(reward), ptr(last_action), ptr(task_ids), ptr(w_r), ptr(b_r),
                                ptr(a_table), a_table.shape[0], ptr(t_table),
                                0 if t_table is None else t_table.shape[0], ptr(out), M, H, stream()),
          "mpx_embed_combine")
    return out


def embed_bwd(dx, reward, last_action, task_ids, A: int, n_tasks: int, d_wr, d_br, d_a_table, d_t_table):
    H = dx.shape[-1]
    M = dx.numel() // H
    check(lib.mpx_embed_bwd(ptr(dx), ptr(reward), ptr(last_action), ptr(task_ids), A, n_tasks, ptr(d_wr), ptr(d_br),
                            ptr(d_a_table), ptr(d_t_table), M, H, stream()), "mpx_embed_bwd")


def policy_logits(x, table, mask=None, out=None):
    H = x.shape[-1]
    M = x.numel() // H
    A = table.shape[0]
    out = torch.empty(M, A, dtype=F32, device=x.device) if out is None else out
    check(lib.mpx_policy_logits(ptr(x), ptr(table), ptr(_u8(mask)), ptr(out), M, H, A, stream()), "mpx_policy_logits")
    return out


def policy_sample(logits, gumbel, action=None, log_prob=None):
    A = logits.shape[-1]
    M = logits.numel() // A
    action = torch.empty(M, dtype=I32, device=logits.device) if action is None else action
    log_prob = torch.empty(M, dtype=F32, device=logits.device) if log_prob is None else log_prob
    check(lib.mpx_policy_sample(ptr(logits), ptr(gumbel), ptr(action), ptr(log_prob), M, A, stream()),
          "mpx_policy_sample")
    return action, log_prob


def policy_head_sample(x, table, action, log_prob, seed: int = 0, stream_id: int = 0, stream_offset=None, mask=None,
                       gumbel_in=None, norm_scale=None, eps: float = 1e-6, logits_out=None, xf_out=None):
    """Fused rollout policy path: (RMSNorm) -> tied logits -> mask -> Gumbel arg-max -> log-prob."""
    _req(x, BF16, "x"); _req(table, F32, "table"); _req(action, I32, "action"); _req(log_prob, F32, "log_prob")
    H = x.shape[-1]
    M = x.numel() // H
    A = table.shape[0]
    check(lib.mpx_policy_head_sample(ptr(x), ptr(norm_scale), eps, ptr(table), ptr(_u8(mask)), ptr(gumbel_in), seed,
                                     stream_id, ptr(stream_offset), ptr(logits_out), ptr(xf_out), ptr(action),
                                     ptr(log_prob), M, H, A, stream()), "mpx_policy_head_sample")
    return action, log_prob


def gumbel_noise(out, seed: int, stream_id: int, stream_offset=None):
    check(lib.mpx_gumbel_noise(ptr(out), out.numel(), seed, stream_id, ptr(stream_offset), stream()),
          "mpx_gumbel_noise")
    return out


def ppo_policy_loss(logits, actions, old_log_prob, advantages, clip_eps: float, entropy_coef: float, losses=None,
                    dlogits=None, want_grad: bool = True, ws=None, entropy_coef_dev=None):
    A = logits.shape[-1]
    M = logits.numel() // A
    losses = torch.empty(2, dtype=F32, device=logits.device) if losses is None else losses
    if dlogits is None and want_grad:
        dlogits = torch.empty_like(logits)
    if ws is None:
        ws = torch.empty(lib.mpx_ppo_policy_loss_workspace_bytes(M), dtype=torch.uint8, device=logits.device)
    check(lib.mpx_ppo_policy_loss(ptr(logits), ptr(actions), ptr(old_log_prob), ptr(advantages), clip_eps,
                                  entropy_coef, ptr(entropy_coef_dev), ptr(losses), ptr(dlogits), ptr(ws), M, A,
                                  stream()),
          "mpx_ppo_policy_loss")
    return losses, dlogits


def policy_head_bwd(dlogits, x, table, dtable, dx_other=None, out=None):
    H = x.shape[-1]
    M = x.numel() // H
    out = torch.empty_like(x) if out is None else out
    check(lib.mpx_policy_head_bwd(ptr(dlogits), ptr(x), ptr(table), ptr(dx_other), ptr(out), ptr(dtable), M, H,
                                  table.shape[0], stream()), "mpx_policy_head_bwd")
    return out


# ------------------------------------------------------------------------------------------------ obs encoder
def obs_onehot_im2col(obs, out, M: int, IH: int, IW: int, sizes: Sequence[int], kh, kw, sh, sw):
    K = len(sizes)
    arr = (C.c_int * K)(*sizes)
    check(lib.mpx_obs_onehot_im2col(ptr(obs), ptr(out), M, IH, IW, K, arr, kh, kw, sh, sw, out.shape[-1], stream()),
          "mpx_obs_onehot_im2col")
    return out


def conv1_onehot_fwd(obs, w, bias, a_out, M: int, IH: int, IW: int, sizes: Sequence[int], kh, kw, sh, sw, cout: int,
                     z_out=None, gelu: bool = True):
    """First conv layer straight from the uint8 observation (no one-hot / im2col materialisation)."""
    K = len(sizes)
    arr = (C.c_int * K)(*sizes)
    check(lib.mpx_conv1_onehot_fwd(ptr(obs), ptr(w), ptr(bias), ptr(z_out), ptr(a_out), M, IH, IW, K, arr, kh, kw, sh,
                                   sw, cout, int(gelu), stream()), "mpx_conv1_onehot_fwd")
    return a_out


def conv1_onehot_wgrad(obs, dz, dw, M: int, IH: int, IW: int, sizes: Sequence[int], kh, kw, sh, sw, cout: int, ws=None):
    """dw (kh*kw*C, cout) f32 += onehot_im2col(obs)^T @ dz without materialising the one-hot rows."""
    _req(dz, BF16, "dz"); _req(dw, F32, "dw")
    K = len(sizes)
    arr = (C.c_int * K)(*sizes)
    need = lib.mpx_conv1_onehot_wgrad_workspace_bytes(M, IH, IW, kh, kw, sh, sw)
    if ws is None:
        ws = wgrad_workspace(dz.device)
    if ws.numel() * ws.element_size() < need:
        raise _lib.MpxError(f"conv1_onehot_wgrad: workspace of {need} bytes required")
    check(lib.mpx_conv1_onehot_wgrad(ptr(obs), ptr(dz), ptr(dw), M, IH, IW, K, arr, kh, kw, sh, sw, cout, ptr(ws),
                                     ws.numel() * ws.element_size(), stream()), "mpx_conv1_onehot_wgrad")
    return dw


def flat_embed_fwd(obs, w, bias, out, M: int, cells: int):
    """FlattenedObsEncoder embedding lookup: out (M, KP) bf16."""
    _req(w, F32, "w"); _req(bias, F32, "bias"); _req(out, BF16, "out")
    check(lib.mpx_flat_embed_fwd(ptr(obs), ptr(w), ptr(bias), ptr(out), M, cells, w.shape[0], out.shape[-1], stream()),
          "mpx_flat_embed_fwd")
    return out


def flat_embed_bwd(obs, dflat, dw, dbias, M: int, cells: int):
    _req(dflat, BF16, "dflat"); _req(dw, F32, "dw"); _req(dbias, F32, "dbias")
    check(lib.mpx_flat_embed_bwd(ptr(obs), ptr(dflat), ptr(dw), ptr(dbias), M, cells, dw.shape[0], dflat.shape[-1],
                                 stream()), "mpx_flat_embed_bwd")


def obs_fused_image(device) -> torch.Tensor:
    """Device buffer for the fused observation encoder's weight image (see mpx_obs_fused_prep)."""
    return torch.empty(lib.mpx_obs_fused_image_bytes(), dtype=torch.uint8, device=device)


def obs_fused_prep(sizes: Sequence[int], w1, b1, wt2, b2, wt3, b3, image):
    _req(w1, BF16, "w1"); _req(wt2, BF16, "wt2"); _req(wt3, BF16, "wt3"); _req(image, torch.uint8, "image")
    K = len(sizes)
    arr = (C.c_int * K)(*sizes)
    check(lib.mpx_obs_fused_prep(arr, K, ptr(w1), ptr(b1), ptr(wt2), ptr(b2), ptr(wt3), ptr(b3), ptr(image), stream()),
          "mpx_obs_fused_prep")
    return image


def obs_embed_fused(obs, sizes: Sequence[int], image, reward, last_action, task_ids, w_r, b_r, a_table, t_table, out,
                    M: int, IH: int, IW: int, obs_emb=None):
    """Whole observation encoder + input embedding sum of a rollout step in one kernel (standard 11x11 geometry)."""
    _req(out, BF16, "out"); _req(image, torch.uint8, "image")
    K = len(sizes)
    arr = (C.c_int * K)(*sizes)
    H = out.shape[-1]
    check(lib.mpx_obs_embed_fused(ptr(obs), arr, K, ptr(image), ptr(reward), ptr(last_action), ptr(task_ids), ptr(w_r),
                                  ptr(b_r), ptr(a_table), a_table.shape[0], ptr(t_table),
                                  0 if t_table is None else t_table.shape[0], ptr(out), ptr(obs_emb), M, IH, IW, H,
                                  stream()), "mpx_obs_embed_fused")
    return out


def obs_train_supported(sizes: Sequence[int]) -> bool:
    K = len(sizes)
    return bool(lib.mpx_obs_train_supported((C.c_int * K)(*sizes), K))


def obs_train_buffers(M: int, device) -> dict:
    """Saved activations / cotangents of the fused training encoder: tile layouts sized for whole 128-token tiles."""
    nt = int(lib.mpx_obs_train_tiles(M))
    e = lambda *s: torch.empty(*s, dtype=BF16, device=device)
    return dict(z1t=e(nt, 25, 2, 128, 8), a1t=e(nt, 25, 2, 128, 8), z2t=e(nt, 9, 4, 128, 8), dz2t=e(nt, 9, 4, 128, 8),
                ws=torch.empty(lib.mpx_obs_train_bwd_workspace_bytes(), dtype=torch.uint8, device=device))


def obs_tile_to_rows(t, M: int):
    """Tile layout [tile][pos][piece][row][8] -> row-major (M, pos, piece*8) (tests / inspection)."""
    nt, P, Q = t.shape[0], t.shape[1], t.shape[2]
    return t.permute(0, 3, 1, 2, 4).reshape(nt * 128, P, Q * 8)[:M]


def obs_train_fwd(obs, sizes: Sequence[int], image, reward, last_action, task_ids, w_r, b_r, a_table, t_table, z1t, a1t,
                  z2t, a2, out, M: int, IH: int, IW: int, obs_emb=None):
    """Training forward of the observation encoder + embedding sum in one kernel; saves z1t, a1t, z2t (tile layouts)
    and a2 (M, 288) for the backward."""
    for t, n in ((out, "out"), (z1t, "z1t"), (a1t, "a1t"), (z2t, "z2t"), (a2, "a2")):
        _req(t, BF16, n)
    _req(image, torch.uint8, "image")
    nt = int(lib.mpx_obs_train_tiles(M))
    if z1t.numel() < nt * 25 * 2048 or a1t.numel() < nt * 25 * 2048 or z2t.numel() < nt * 9 * 4096:
        raise _lib.MpxError("obs_train_fwd: tile-layout buffers too small (see obs_train_buffers)")
    K = len(sizes)
    arr = (C.c_int * K)(*sizes)
    check(lib.mpx_obs_train_fwd(ptr(obs), arr, K, ptr(image), ptr(reward), ptr(last_action), ptr(task_ids), ptr(w_r),
                                ptr(b_r), ptr(a_table), a_table.shape[0], ptr(t_table),
                                0 if t_table is None else t_table.shape[0], ptr(z1t), ptr(a1t), ptr(z2t), ptr(a2), ptr(out),
                                ptr(obs_emb), M, IH, IW, out.shape[-1], stream()), "mpx_obs_train_fwd")
    return out


def obs_train_bwd(dx0, z1t, z2t, image, dz2t, dz1, db1, db2, M: int, ws):
    """Data gradients of the fused training encoder: dz2t (tile layout), dz1 (M*25, 16); db1 / db2 (fp32) +=."""
    _req(dx0, BF16, "dx0"); _req(dz1, BF16, "dz1"); _req(db1, F32, "db1"); _req(db2, F32, "db2")
    nt = int(lib.mpx_obs_train_tiles(M))
    if z1t.numel() < nt * 25 * 2048 or z2t.numel() < nt * 9 * 4096 or dz2t.numel() < nt * 9 * 4096 or dz1.numel() < M * 400:
        raise _lib.MpxError("obs_train_bwd: buffers too small")
    check(lib.mpx_obs_train_bwd(ptr(dx0), ptr(z1t), ptr(z2t), ptr(image), ptr(dz2t), ptr(dz1), ptr(db1), ptr(db2), M,
                                ptr(ws), ws.numel() * ws.element_size(), stream()), "mpx_obs_train_bwd")


def obs_conv2_wgrad(a1t, dz2t, dw, M: int, ws):
    """dw (144, 32) f32 += tap-wise conv2 weight gradient from the position tiles (no im2col)."""
    _req(dw, F32, "dw")
    if dw.numel() != 144 * 32 or not dw.is_contiguous():
        raise _lib.MpxError("obs_conv2_wgrad: dw must be the contiguous (3,3,16,32) kernel gradient")
    check(lib.mpx_obs_conv2_wgrad(ptr(a1t), ptr(dz2t), ptr(dw), M, ptr(ws), ws.numel() * ws.element_size(), stream()),
          "mpx_obs_conv2_wgrad")


def im2col(a, out, M: int, IH: int, IW: int, Cc: int, kh, kw, sh, sw):
    check(lib.mpx_im2col(ptr(a), ptr(out), M, IH, IW, Cc, kh, kw, sh, sw, stream()), "mpx_im2col")
    return out


def col2im(dx, out, M: int, IH: int, IW: int, Cc: int, kh, kw, sh, sw, z=None):
    """Transpose of im2col; with z also the gelu transpose of the layer below: out = bf16(bf16(sum) * gelu'(z))."""
    check(lib.mpx_col2im(ptr(dx), ptr(out), M, IH, IW, Cc, kh, kw, sh, sw, ptr(z), stream()), "mpx_col2im")
    return out


# ------------------------------------------------------------------------------------------------ optimizer etc.
def adamw_step(params, grads, mu, nu, scratch, step, lr0, total_steps, b1, b2, eps, wd, max_norm, grad_scale=1.0,
               norm_out=None, ws=None):
    n = params.numel()
    if ws is None:
        ws = torch.empty(lib.mpx_adamw_workspace_bytes(n), dtype=torch.uint8, device=params.device)
    check(lib.mpx_adamw_step(ptr(params), ptr(grads), ptr(mu), ptr(nu), ptr(scratch), n, ptr(step), lr0,
                             float(total_steps), b1, b2, eps, wd, max_norm, grad_scale, ptr(norm_out), ptr(ws),
                             stream()), "mpx_adamw_step")


def muon_step(params, grads, mu, nu, scratch, n_adam: int, table, nmat: int, x_total: int, a_total: int, max_r: int,
              max_c: int, ns_ws, step, lr0, total_steps, b1, b2, eps, wd, max_norm, grad_scale=1.0, muon_beta=0.95,
              ns_steps=5, norm_out=None, ws=None, phases: int = 3):
    """optax.contrib.muon partition (optimizer.py:25-37): Muon on the labelled matrices, AdamW(nesterov) elsewhere.
    phases: 1 = matrices only (enqueue early, e.g. on a side stream), 2 = AdamW leaves + clip + apply, 3 = both."""
    n = params.numel()
    if ws is None:
        if phases != 3:
            raise _lib.MpxError("muon_step: the two phases must share one caller-provided workspace")
        ws = torch.empty(lib.mpx_muon_workspace_bytes(n, nmat), dtype=torch.uint8, device=params.device)
    check(lib.mpx_muon_step_phased(ptr(params), ptr(grads), ptr(mu), ptr(nu), ptr(scratch), n, n_adam, ptr(table), nmat,
                                   x_total, a_total, max_r, max_c, ptr(ns_ws), ptr(step), lr0, float(total_steps), b1, b2,
                                   eps, wd, muon_beta, ns_steps, max_norm, grad_scale, ptr(norm_out), ptr(ws), phases,
                                   stream()), f"mpx_muon_step:{phases}")


def muon_apply_sharded(params, grads, mu, nu, scratch, n_adam: int, step, lr0, total_steps, b1, b2, eps, wd, max_norm,
                       grad_scale=1.0, norm_out=None, ws=None):
    """Second half of the sharded Muon step: AdamW leaves + global-norm clip over the all-gathered update + apply."""
    n = params.numel()
    check(lib.mpx_muon_apply_sharded(ptr(params), ptr(grads), ptr(mu), ptr(nu), ptr(scratch), n, n_adam, ptr(step), lr0,
                                     float(total_steps), b1, b2, eps, wd, max_norm, grad_scale, ptr(norm_out), ptr(ws),
                                     stream()), "mpx_muon_apply_sharded")


def prep_weights(descs_device, ndesc: int):
    check(lib.mpx_prep_weights(ptr(descs_device), ndesc, stream()), "mpx_prep_weights")


def transpose_tb(src, dst, perm, T: int, B: int):
    item_bytes = src.numel() * src.element_size() // (T * B)
    check(lib.mpx_transpose_tb(ptr(src), ptr(dst), ptr(perm), T, B, item_bytes, stream()), "mpx_transpose_tb")
    return dst
