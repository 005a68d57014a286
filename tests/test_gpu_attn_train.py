"""Training attention forward / backward against the oracle (autograd through the XLA-path restatement)."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from _util import ULP2, assert_close_bf16, log as _log, rope_timescale as _rope_timescale  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("B,T,Nq,Nkv", [(2, 64, 4, 1), (3, 128, 4, 1), (1, 512, 4, 1), (5, 256, 4, 1), (2, 128, 4, 2),
                                        (2, 64, 4, 4)])
def test_attn_train_fwd_bwd(B, T, Nq, Nkv, impl):
    """impl 0: tcgen05/TMEM forward (four heads per CTA when Nq == 4) AND backward where the shape allows (Nkv == 1,
    T % 128 == 0), impl 1: mma.sync, impl 2: tcgen05 forward with one head per CTA."""
    from jaxrl_b200 import ops
    from jaxrl_b200._lib import lib
    lib.mpx_debug_set(2, impl)
    D = 32
    gen = torch.Generator().manual_seed(B * 100 + T + Nq + Nkv)
    QD, KD = Nq * D, Nkv * D
    QKV = QD + 2 * KD
    pos = torch.arange(T, dtype=torch.int32)
    # un-rotated projections (what the q/k/v Linear layers output), rotated here by the oracle
    q0 = torch.randn(B, T, Nq, D, generator=gen).to(BF16).float().requires_grad_(True)
    k0 = torch.randn(B, T, Nkv, D, generator=gen).to(BF16).float().requires_grad_(True)
    v0 = torch.randn(B, T, Nkv, D, generator=gen).to(BF16).float().requires_grad_(True)
    q = O.apply_rope(q0.to(BF16), pos[None], D)
    k = O.apply_rope(k0.to(BF16), pos[None], D)
    v = v0.to(BF16)
    ref = O.dot_product_attention(q, k, v, is_causal=True)
    dout = torch.randn(B, T, Nq, D, generator=gen).to(BF16)
    (ref.float() * dout.float()).sum().backward()

    qkv = torch.cat([q.reshape(B * T, QD), k.reshape(B * T, KD), v.reshape(B * T, KD)], dim=1).detach().to(DEV).contiguous()
    dqkv = torch.zeros(B * T, QKV, dtype=BF16, device=DEV)
    try:
        o, lse = ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], B, T, Nq, Nkv, D, QKV, QKV, QKV)
        torch.cuda.synchronize()
        ops.attn_train_bwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], o, dout.reshape(B * T, QD).to(DEV).contiguous(), lse,
                           pos.to(DEV), _rope_timescale(D).to(DEV), B, T, Nq, Nkv, D, QKV, QKV, QKV,
                           dqkv, dqkv[:, QD:], dqkv[:, QD + KD:], QKV, QKV, QKV)
        torch.cuda.synchronize()
    finally:
        lib.mpx_debug_set(2, 0)
    # flash-style online softmax rounds the UN-normalised probabilities to bf16 (as cuDNN's fused SDPA, the
    # reference's GPU implementation, does); the XLA path rounds after normalising -> up to 2 bf16 ulps apart.
    assert_close_bf16(o.reshape(B, T, Nq, D), ref, f"attn_train_fwd impl{impl} B{B} T{T} kv{Nkv}", rtol=ULP2, abs_floor_frac=1e-2,
                      allow_frac=2e-4, max_err_frac=0.15)
    got = dqkv.float().cpu()
    for name, g, ref_g in (("dq", got[:, :QD], q0.grad.reshape(B * T, QD)),
                           ("dk", got[:, QD:QD + KD], k0.grad.reshape(B * T, KD)),
                           ("dv", got[:, QD + KD:], v0.grad.reshape(B * T, KD))):
        rel = (g - ref_g).norm().item() / (ref_g.norm().item() + 1e-9)
        _log(f"attn_train_bwd impl{impl} B{B} T{T} kv{Nkv} {name}: rel_l2={rel:.4e}")
        assert rel < 8e-3, f"{name} rel l2 error {rel}"
