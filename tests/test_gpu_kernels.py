"""GPU parity tests: every C-ABI kernel against the CPU oracle on identical seeded inputs.

Tolerances follow BASELINE.json: bf16 results <= 1e-2 relative, fp32 GAE <= 1e-5, indices bit-exact.
"""

import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from _util import ULP2, assert_close_bf16, log as _log, rope_timescale as _rope_timescale  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from jaxrl_b200 import ops as _ops
    return _ops


DEV = "cuda"
BF16 = torch.bfloat16


def g(seed):
    return torch.Generator().manual_seed(seed)


# ---------------------------------------------------------------------------------------------- GAE
@pytest.mark.parametrize("B,T", [(1, 1), (2, 4), (37, 100), (256, 512), (4096, 512)])
def test_gae_matches_oracle(ops, B, T):
    gen = g(B * 1000 + T)
    r = torch.randn(B, T, generator=gen)
    v = torch.randn(B, T + 1, generator=gen)
    term = torch.rand(B, T, generator=gen) < (1.0 / 16)
    adv_ref, tgt_ref = O.calculate_advantage(r.numpy(), v.numpy(), term.numpy(), 0.9660902557298033,
                                             0.9123716252262187, False)
    adv, tgt, stats = ops.gae(r.to(DEV), v.to(DEV), term.to(DEV), 0.9660902557298033, 0.9123716252262187)
    np.testing.assert_array_equal(tgt.cpu().numpy(), tgt_ref)        # sequential fp32: bit exact
    np.testing.assert_array_equal(adv.cpu().numpy(), adv_ref)
    assert abs(stats[0].item() - adv_ref.astype(np.float64).sum()) <= 1e-9 * max(1.0, abs(stats[0].item())) + 1e-6
    # normalisation
    adv_n_ref, _ = O.calculate_advantage(r.numpy(), v.numpy(), term.numpy(), 0.9660902557298033,
                                         0.9123716252262187, True)
    if B * T > 1:
        ops.adv_normalize(adv, stats)
        np.testing.assert_allclose(adv.cpu().numpy(), adv_n_ref, rtol=1e-5, atol=1e-5)


def test_gae_known_answers(ops):
    # reference tests/test_advantage.py:39-61 and :150-163
    adv, tgt, _ = ops.gae(torch.tensor([[1.0]], device=DEV), torch.tensor([[0.5, 0.8]], device=DEV),
                          torch.tensor([[False]], device=DEV), 0.99, 0.95)
    assert abs(tgt.item() - (1.0 + 0.99 * 0.8)) < 1e-5 and abs(adv.item() - (1.0 + 0.99 * 0.8 - 0.5)) < 1e-5
    r = torch.tensor([[5.0, 10.0, 15.0]], device=DEV)
    _, tgt, _ = ops.gae(r, torch.tensor([[1.0, 2.0, 3.0, 4.0]], device=DEV),
                        torch.zeros(1, 3, dtype=torch.bool, device=DEV), 0.0, 0.95)
    assert torch.allclose(tgt, r, atol=1e-5)


# ---------------------------------------------------------------------------------------------- HL-Gauss
@pytest.mark.parametrize("n,NL,vmin,vmax,sigma", [(7, 51, -10.0, 10.0, 0.1038127), (5000, 51, -10.0, 10.0, 0.1038127),
                                                  (300, 51, -5.0, 5.0, 0.75), (64, 10, -1.0, 1.0, 0.5),
                                                  (200, 100, -10.0, 10.0, 0.3), (33, 64, -2.0, 2.0, 0.2)])
def test_hlgauss(ops, n, NL, vmin, vmax, sigma):
    gen = g(n + NL)
    logits = torch.randn(n, NL, generator=gen) * 2.0
    targets = torch.randn(n, generator=gen) * (vmax - vmin) * 0.4
    targets[0] = 100.0
    targets[-1] = -100.0     # clipped (tests/test_values.py:92-98)
    supports, centers = O.calculate_supports(vmin, vmax, NL)
    val_ref = O.hlgauss_get_value(logits, centers)
    val = ops.hlgauss_value(logits.to(DEV), centers.to(DEV))
    torch.testing.assert_close(val.cpu(), val_ref, rtol=1e-5, atol=1e-5)
    lg = logits.clone().requires_grad_(True)
    loss_ref = O.hlgauss_get_loss(lg[None], targets[None], vmin, vmax, sigma)
    loss_ref.backward()
    loss, dl = ops.hlgauss_loss(logits.to(DEV), targets.to(DEV), supports.reshape(-1).to(DEV), vmin, vmax, sigma, 1.0)
    torch.testing.assert_close(loss.cpu()[0], loss_ref.detach(), rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(dl.cpu(), lg.grad, rtol=1e-4, atol=1e-7)
    assert torch.isfinite(dl).all()


def test_hlgauss_uniform_logits_value_zero(ops):
    # tests/test_values.py:49-57
    _, centers = O.calculate_supports(-5.0, 5.0, 51)
    v = ops.hlgauss_value(torch.zeros(2, 4, 51, device=DEV), centers.to(DEV))
    assert v.shape == (2, 4) and torch.allclose(v, torch.zeros_like(v), atol=1e-5)


# ---------------------------------------------------------------------------------------------- dense
def _lin_ref(x, w_t, bias=None, residual=None, act=0):
    y = O.linear(x, w_t.t().float(), None if bias is None else bias.float())
    if act == 1:
        y = O.gelu_tanh(y)
    if residual is not None:
        y = (y.float() + residual.float()).to(BF16)
    return y


@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (300, 128, 192), (4096, 128, 128), (1000, 768, 128),
                                   (4096, 128, 1536), (20000, 128, 768), (77, 128, 51), (513, 192, 128)])
def test_linear_plain(ops, M, K, N):
    gen = g(M + K + N)
    x = (torch.randn(M, K, generator=gen)).to(BF16)
    wt = (torch.randn(N, K, generator=gen) / math.sqrt(K)).to(BF16)
    y = ops.linear(x.to(DEV), wt.to(DEV))
    assert_close_bf16(y, _lin_ref(x, wt), f"linear {M}x{K}x{N}")


def test_linear_epilogues(ops):
    gen = g(11)
    M, K, N = 1000, 128, 768
    x = torch.randn(M, K, generator=gen).to(BF16)
    wt = (torch.randn(N, K, generator=gen) / math.sqrt(K)).to(BF16)
    bias = (torch.randn(N, generator=gen) * 0.5).to(BF16)
    y = ops.linear(x.to(DEV), wt.to(DEV), bias=bias.to(DEV), act=1)
    assert_close_bf16(y, _lin_ref(x, wt, bias=bias, act=1), "linear bias+gelu")
    M, K, N = 777, 768, 128
    x = torch.randn(M, K, generator=gen).to(BF16)
    wt = (torch.randn(N, K, generator=gen) / math.sqrt(K)).to(BF16)
    res = torch.randn(M, N, generator=gen).to(BF16)
    y = ops.linear(x.to(DEV), wt.to(DEV), residual=res.to(DEV))
    assert_close_bf16(y, _lin_ref(x, wt, residual=res), "linear residual")
    out32 = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.linear(x.to(DEV), wt.to(DEV), out_f32=out32, want_bf16=False)
    torch.testing.assert_close(out32.cpu(), x.float() @ wt.float().t(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("M,K,N", [(1000, 64, 768), (777, 128, 288), (333, 128, 96), (4096, 64, 768), (130, 128, 136)])
def test_linear_gelu_transpose_epilogue(ops, M, K, N):
    """act=2: dz = bf16(bf16(dy @ W) * gelu'(z)) in the GEMM epilogue == mpx_linear followed by mpx_gelu_bwd, bit for
    bit (wide tiles fetch z one chunk ahead; narrow tiles before the accumulator wait), and a residual add with a
    wide tile (the same prefetch path)."""
    gen = g(M + N)
    x = torch.randn(M, K, generator=gen).to(BF16).to(DEV)
    wt = (torch.randn(N, K, generator=gen) / math.sqrt(K)).to(BF16).to(DEV)
    z = (torch.randn(M, N, generator=gen) * 1.5).to(BF16).to(DEV)
    two = ops.gelu_bwd(ops.linear(x, wt), z)
    one = ops.linear(x, wt, residual=z, act=2)
    assert torch.equal(one, two), "fused gelu transpose differs from linear + gelu_bwd"
    y = ops.linear(x, wt, residual=z)
    assert_close_bf16(y, _lin_ref(x.cpu(), wt.cpu(), residual=z.cpu()), "linear residual (wide tile)")


@pytest.mark.parametrize("M,K,N", [(64, 128, 64), (1000, 128, 192), (131072, 128, 1536), (5000, 768, 128), (333, 128, 128)])
def test_linear_wgrad(ops, M, K, N):
    gen = g(M + 3 * K + N)
    x = torch.randn(M, K, generator=gen).to(BF16)
    dy = (torch.randn(M, N, generator=gen) * 0.1).to(BF16)
    ref = (x.double().t() @ dy.double()).float() if M <= 5000 else None
    xd, dyd = x.to(DEV), dy.to(DEV)
    if ref is None:
        ref = (xd.float().t() @ dyd.float()).cpu()
    from jaxrl_b200._lib import lib
    results = {}
    for variant in (0, 1):
        lib.mpx_debug_set(0, variant)
        dw = torch.zeros(K, N, dtype=torch.float32, device=DEV)
        ops.linear_wgrad(xd, dyd, dw)
        torch.cuda.synchronize()
        err = (dw.cpu() - ref).abs().max().item()
        results[variant] = err
    lib.mpx_debug_set(0, 0)
    _log(f"wgrad {M}x{K}x{N}: max_abs_err variant0={results[0]:.4e} variant1(swapped)={results[1]:.4e} "
         f"ref_rms={ref.pow(2).mean().sqrt().item():.4e}")
    tol = 2e-3 * ref.pow(2).mean().sqrt().item() * math.sqrt(max(M, 64) / 64)
    assert results[0] <= max(tol, 1e-3), f"wgrad variant0 err {results[0]} (variant1 {results[1]})"


@pytest.mark.parametrize("Nq,Nkv", [(4, 1), (4, 2), (4, 4)])
def test_qkv_rope_train(ops, Nq, Nkv):
    gen = g(5 + Nq + Nkv)
    Bb, T, H, D = 3, 40, 128, 32
    x = torch.randn(Bb, T, H, generator=gen).to(BF16)
    wq = torch.randn(H, Nq * D, generator=gen) / math.sqrt(H)
    wk = torch.randn(H, Nkv * D, generator=gen) / math.sqrt(H)
    wv = torch.randn(H, Nkv * D, generator=gen) / math.sqrt(H)
    pos = torch.arange(T, dtype=torch.int32)
    q_ref = O.apply_rope(O.linear(x, wq).reshape(Bb, T, Nq, D), pos[None], D)
    k_ref = O.apply_rope(O.linear(x, wk).reshape(Bb, T, Nkv, D), pos[None], D)
    v_ref = O.linear(x, wv).reshape(Bb, T, Nkv, D)
    wt = torch.cat([wq, wk, wv], dim=1).t().contiguous().to(BF16)
    q, k, v = ops.qkv_rope(x.reshape(-1, H).to(DEV), wt.to(DEV), pos.to(DEV), _rope_timescale(D).to(DEV), Nq, Nkv, D)
    assert_close_bf16(q.reshape(Bb, T, Nq, D), q_ref, "qkv_rope q")
    assert_close_bf16(k.reshape(Bb, T, Nkv, D), k_ref, "qkv_rope k")
    assert_close_bf16(v.reshape(Bb, T, Nkv, D), v_ref, "qkv_rope v")


@pytest.mark.parametrize("step,S", [(0, 16), (3, 8), (6, 4), (511, 512)])
def test_qkv_rope_cache_slot(ops, step, S):
    """Ring write at pos % S (reference tests/test_attention.py:91-121), untouched elsewhere (bit exact)."""
    gen = g(step + S)
    Bb, H, D, Nq, Nkv = 5, 128, 32, 4, 1
    x = torch.randn(Bb, H, generator=gen).to(BF16)
    w = (torch.randn((Nq + 2 * Nkv) * D, H, generator=gen) / math.sqrt(H)).to(BF16)
    kc = torch.full((Bb, S, Nkv, D), 7.0, dtype=BF16, device=DEV)
    vc = torch.full((Bb, S, Nkv, D), -3.0, dtype=BF16, device=DEV)
    pos = torch.full((Bb,), step, dtype=torch.int32, device=DEV)
    q, _, _ = ops.qkv_rope(x.to(DEV), w.to(DEV), pos, _rope_timescale(D).to(DEV), Nq, Nkv, D, k=kc, v=vc,
                           k_row_stride=S * Nkv * D, v_row_stride=S * Nkv * D, cache_S=S)
    slot = step % S
    k_ref = O.apply_rope(O.linear(x, w[Nq * D:(Nq + Nkv) * D].t().float()).reshape(Bb, 1, Nkv, D),
                         torch.full((Bb, 1), step), D)
    v_ref = O.linear(x, w[(Nq + Nkv) * D:].t().float()).reshape(Bb, 1, Nkv, D)
    assert_close_bf16(kc[:, slot], k_ref[:, 0], "cache k slot")
    assert_close_bf16(vc[:, slot], v_ref[:, 0], "cache v slot")
    mask = torch.ones(S, dtype=torch.bool)
    mask[slot] = False
    assert (kc[:, mask] == 7.0).all() and (vc[:, mask] == -3.0).all()


@pytest.mark.parametrize("Bb", [5, 300, 4096])
def test_qkv_rope_fused_prenorm(ops, Bb):
    """history_norm (network.py:160) folded into the q/k/v GEMM: same result as rms_norm -> qkv_rope."""
    gen = g(Bb)
    H, D, Nq, Nkv, S, step = 128, 32, 4, 1, 16, 5
    x = (torch.randn(Bb, H, generator=gen) * 3).to(BF16)
    scale = 1.0 + 0.2 * torch.randn(H, generator=gen)
    w = (torch.randn((Nq + 2 * Nkv) * D, H, generator=gen) / math.sqrt(H)).to(BF16)
    kc = torch.zeros((Bb, S, Nkv, D), dtype=BF16, device=DEV)
    vc = torch.zeros((Bb, S, Nkv, D), dtype=BF16, device=DEV)
    pos = torch.full((Bb,), step, dtype=torch.int32, device=DEV)
    q, _, _ = ops.qkv_rope(x.to(DEV), w.to(DEV), pos, _rope_timescale(D).to(DEV), Nq, Nkv, D, k=kc, v=vc,
                           k_row_stride=S * Nkv * D, v_row_stride=S * Nkv * D, cache_S=S, norm_scale=scale.to(DEV))
    xn = O.rms_norm(x, scale)
    posr = torch.full((Bb, 1), step)
    q_ref = O.apply_rope(O.linear(xn, w[:Nq * D].t().float()).reshape(Bb, 1, Nq, D), posr, D)
    k_ref = O.apply_rope(O.linear(xn, w[Nq * D:(Nq + Nkv) * D].t().float()).reshape(Bb, 1, Nkv, D), posr, D)
    v_ref = O.linear(xn, w[(Nq + Nkv) * D:].t().float()).reshape(Bb, 1, Nkv, D)
    assert_close_bf16(q.reshape(Bb, 1, Nq, D), q_ref, "prenorm q")
    assert_close_bf16(kc[:, step % S], k_ref[:, 0], "prenorm k")
    assert_close_bf16(vc[:, step % S], v_ref[:, 0], "prenorm v")


def test_glu_up(ops):
    gen = g(21)
    M, K, F = 700, 128, 768
    x = torch.randn(M, K, generator=gen).to(BF16)
    wu = torch.randn(K, F, generator=gen) / math.sqrt(K)
    wg = torch.randn(K, F, generator=gen) / math.sqrt(K)
    up = O.linear(x, wu)
    gate = O.linear(x, wg)
    h_ref = (O.gelu_tanh(up).float() * gate.float()).to(BF16)
    h, u, gg = ops.glu_up(x.to(DEV), ops.pack_glu_weight(wu, wg).to(DEV), F, save=True)
    assert_close_bf16(u, up, "glu u")
    assert_close_bf16(gg, gate, "glu g")
    assert_close_bf16(h, h_ref, "glu h", rtol=2e-2)


# ---------------------------------------------------------------------------------------------- decode attention
@pytest.mark.parametrize("step,S,Nq,Nkv,Bb", [(0, 16, 4, 1, 3), (5, 16, 4, 1, 9), (31, 64, 4, 1, 4), (32, 64, 4, 1, 4),
                                              (100, 512, 4, 1, 33), (511, 512, 4, 1, 130), (700, 512, 4, 1, 5),
                                              (77, 128, 4, 2, 6), (40, 64, 4, 4, 3), (9, 32, 8, 1, 2)])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_attn_decode(ops, step, S, Nq, Nkv, Bb, impl):
    """impl 0 = mma.sync online softmax (default), 1 = SIMT exact two-pass, 2 = mma.sync exact two-pass."""
    from jaxrl_b200._lib import lib
    lib.mpx_debug_set(1, impl)
    gen = g(step * 7 + S + Nq + Nkv)
    D = 32
    q = torch.randn(Bb, Nq, D, generator=gen).to(BF16)
    kc = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16)
    vc = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16)
    kv_len = min(step + 1, S)
    ref = O.dot_product_attention(q[:, None], kc, vc, kv_len=kv_len)[:, 0]
    pos = torch.full((Bb,), step, dtype=torch.int32, device=DEV)
    try:
        out = ops.attn_decode(q.reshape(Bb, -1).to(DEV), kc.to(DEV), vc.to(DEV), pos, Nq, Nkv, D)
        torch.cuda.synchronize()
    finally:
        lib.mpx_debug_set(1, 0)
    # the online-softmax kernel rounds un-normalised probabilities (<= 2 bf16 ulps from the XLA arithmetic)
    assert_close_bf16(out.reshape(Bb, Nq, D), ref, f"attn_decode impl={impl} step={step} S={S} Nkv={Nkv}",
                      rtol=ULP2 if impl == 0 else 1e-2, abs_floor_frac=1e-2 if impl == 0 else 4e-3)


@pytest.mark.parametrize("step,S,Bb", [(0, 16, 3), (5, 16, 9), (31, 64, 14), (32, 64, 15), (100, 512, 33), (300, 512, 4096),
                                       (511, 512, 130), (700, 512, 29), (1029, 512, 2100)])
@pytest.mark.parametrize("H", [128, 256])
def test_attn_decode_fused(ops, step, S, Bb, H):
    """history_norm + q/k/v + RoPE + ring write + attention in one kernel vs the oracle chain (network.py:158-165,
    attention.py:96-150); the ring must change only at slot pos % S, there by exactly the new k,v.  H = 256 is the
    multitask width (config/multitask.json:31)."""
    gen = g(step * 3 + S + Bb + H)
    D, Nq, Nkv = 32, 4, 1
    x = (torch.randn(Bb, H, generator=gen) * 2).to(BF16)
    scale = 1.0 + 0.2 * torch.randn(H, generator=gen)
    w = (torch.randn((Nq + 2 * Nkv) * D, H, generator=gen) / math.sqrt(H)).to(BF16)
    kc0 = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16)
    vc0 = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16)
    kc, vc = kc0.clone().to(DEV), vc0.clone().to(DEV)
    pos = torch.full((Bb,), step, dtype=torch.int32, device=DEV)
    out = ops.attn_decode_fused(x.to(DEV), scale.to(DEV), w.to(DEV), pos, _rope_timescale(D).to(DEV), kc, vc, Nq, Nkv, D)
    torch.cuda.synchronize()
    nb = min(Bb, 256)                      # oracle on the first and last agents
    sel = torch.cat([torch.arange(nb // 2), torch.arange(Bb - (nb - nb // 2), Bb)]) if Bb > nb else torch.arange(Bb)
    xn = O.rms_norm(x[sel], scale)
    posr = torch.full((len(sel), 1), step)
    q_ref = O.apply_rope(O.linear(xn, w[:Nq * D].t().float()).reshape(len(sel), 1, Nq, D), posr, D)
    k_ref = O.apply_rope(O.linear(xn, w[Nq * D:(Nq + Nkv) * D].t().float()).reshape(len(sel), 1, Nkv, D), posr, D)
    v_ref = O.linear(xn, w[(Nq + Nkv) * D:].t().float()).reshape(len(sel), 1, Nkv, D)
    slot = step % S
    kr, vr = kc0[sel].clone(), vc0[sel].clone()
    kr[:, slot] = k_ref[:, 0]
    vr[:, slot] = v_ref[:, 0]
    ref = O.dot_product_attention(q_ref, kr, vr, kv_len=min(step + 1, S))[:, 0]
    assert_close_bf16(kc[sel.to(DEV), slot], k_ref[:, 0], "fused decode: new k")
    assert_close_bf16(vc[sel.to(DEV), slot], v_ref[:, 0], "fused decode: new v")
    mask = torch.ones(S, dtype=torch.bool)
    mask[slot] = False
    assert torch.equal(kc.cpu()[:, mask], kc0[:, mask]) and torch.equal(vc.cpu()[:, mask], vc0[:, mask])
    assert_close_bf16(out[sel.to(DEV)].reshape(len(sel), Nq, D), ref, f"attn_decode_fused step={step} S={S} B={Bb}", rtol=ULP2,
                      abs_floor_frac=1e-2)
    # and against the three-kernel path on the device, all agents
    kc2, vc2 = kc0.clone().to(DEV), vc0.clone().to(DEV)
    xn_d = ops.rmsnorm_fwd(x.to(DEV), scale.to(DEV))
    q_d, _, _ = ops.qkv_rope(xn_d, w.to(DEV), pos, _rope_timescale(D).to(DEV), Nq, Nkv, D, k=kc2, v=vc2,
                             k_row_stride=S * Nkv * D, v_row_stride=S * Nkv * D, cache_S=S)
    out2 = ops.attn_decode(q_d, kc2, vc2, pos, Nq, Nkv, D)
    torch.cuda.synchronize()
    assert_close_bf16(out, out2.cpu(), "attn_decode_fused vs rmsnorm + qkv_rope + attn_decode", rtol=1e-2, abs_floor_frac=1e-2)


@pytest.mark.parametrize("step,S,Bb", [(0, 16, 3), (40, 64, 15), (300, 512, 4096), (511, 512, 130), (1029, 512, 2100)])
@pytest.mark.parametrize("H", [128, 256])
def test_attn_decode_fused_early_stream(ops, step, S, Bb, H):
    """flags bit 0 (ring history declared stable: its stream starts before griddepcontrol.wait) changes neither the
    output nor the ring, bit for bit, also right behind a kernel that has just rewritten x."""
    gen = g(step + S + Bb + H)
    D, Nq, Nkv = 32, 4, 1
    x = (torch.randn(Bb, H, generator=gen) * 2).to(BF16).to(DEV)
    scale = (1.0 + 0.2 * torch.randn(H, generator=gen)).to(DEV)
    w = (torch.randn((Nq + 2 * Nkv) * D, H, generator=gen) / math.sqrt(H)).to(BF16).to(DEV)
    kc0 = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16).to(DEV)
    vc0 = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16).to(DEV)
    pos = torch.full((Bb,), step, dtype=torch.int32, device=DEV)
    ts = _rope_timescale(D).to(DEV)
    res = []
    for early in (False, True):
        kc, vc = kc0.clone(), vc0.clone()
        xin = torch.zeros_like(x)
        xin.copy_(x)                                     # the producer of x is the kernel right before the launch
        out = ops.attn_decode_fused(xin, scale, w, pos, ts, kc, vc, Nq, Nkv, D, stable_history=early)
        torch.cuda.synchronize()
        res.append((out.clone(), kc, vc))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_attn_decode_full_batch(ops):
    """Full-size shape (B=4096, S=512): persistent-warp striding covers every agent; bit-for-bit vs the SIMT kernel."""
    from jaxrl_b200._lib import lib
    gen = g(99)
    Bb, S, Nq, Nkv, D, step = 4096, 512, 4, 1, 32, 300
    q = torch.randn(Bb, Nq * D, generator=gen).to(BF16).to(DEV)
    kc = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16).to(DEV)
    vc = torch.randn(Bb, S, Nkv, D, generator=gen).to(BF16).to(DEV)
    pos = torch.full((Bb,), step, dtype=torch.int32, device=DEV)
    outs = []
    for impl in (0, 1, 2):
        lib.mpx_debug_set(1, impl)
        outs.append(ops.attn_decode(q, kc, vc, pos, Nq, Nkv, D).float().cpu())
    lib.mpx_debug_set(1, 0)
    ref = O.dot_product_attention(q[:64].cpu().reshape(64, 1, Nq, D), kc[:64].cpu(), vc[:64].cpu(), kv_len=step + 1)[:, 0]
    assert_close_bf16(outs[2][:64].reshape(64, Nq, D), ref, "attn_decode (exact mma) full batch vs oracle (first 64 agents)")
    assert_close_bf16(outs[2], outs[1], "attn_decode exact mma vs simt (all 4096 agents)", rtol=1e-2)
    assert_close_bf16(outs[0], outs[1], "attn_decode online mma vs simt (all 4096 agents)", rtol=ULP2, abs_floor_frac=1e-2,
                      allow_frac=1e-4, max_err_frac=0.1)


def test_conv1_onehot_matches_oracle_conv(ops):
    gen = g(55)
    M, IH, IW, sizes, cout = 37, 11, 11, (5, 5), 16
    obs = torch.stack([torch.randint(0, n, (M, IH, IW), generator=gen) for n in sizes], -1).to(torch.uint8)
    W = torch.randn(3, 3, sum(sizes), cout, generator=gen) * 0.2
    b = torch.randn(cout, generator=gen) * 0.1
    x = O.concat_one_hot(obs, sizes)
    z_ref = O.conv_valid(x, W, b, (2, 2))
    a_ref = O.gelu_tanh(z_ref)
    z = torch.empty(M * 25, cout, dtype=BF16, device=DEV)
    a = torch.empty(M * 25, cout, dtype=BF16, device=DEV)
    ops.conv1_onehot_fwd(obs.to(DEV), W.reshape(-1, cout).to(BF16).to(DEV), b.to(BF16).to(DEV), a, M, IH, IW, sizes,
                         3, 3, 2, 2, cout, z_out=z)
    assert_close_bf16(z.reshape(M, 5, 5, cout), z_ref, "conv1 onehot z")
    assert_close_bf16(a.reshape(M, 5, 5, cout), a_ref, "conv1 onehot gelu", rtol=2e-2)
    # the one-hot im2col rows (used for the weight gradient) reproduce the same conv through the GEMM
    KP = 96
    X1 = torch.empty(M * 25, KP, dtype=BF16, device=DEV)
    ops.obs_onehot_im2col(obs.to(DEV), X1, M, IH, IW, sizes, 3, 3, 2, 2)
    wt = torch.zeros(cout, KP)
    wt[:, :90] = W.reshape(90, cout).t()
    y = ops.linear(X1, wt.to(BF16).to(DEV), bias=b.to(BF16).to(DEV))
    assert_close_bf16(y.reshape(M, 5, 5, cout), z_ref, "onehot_im2col + GEMM == conv1")
    assert (X1.float().sum(-1) == 18).all()        # 9 taps x 2 components set exactly one channel each


# ---------------------------------------------------------------------------------------------- row-wise / heads
@pytest.mark.parametrize("M,H", [(5, 128), (1000, 128), (777, 256)])
def test_rmsnorm_fwd_bwd(ops, M, H):
    gen = g(M + H)
    x = torch.randn(M, H, generator=gen).to(BF16)
    scale = 1.0 + 0.1 * torch.randn(H, generator=gen)
    y = ops.rmsnorm_fwd(x.to(DEV), scale.to(DEV))
    assert_close_bf16(y, O.rms_norm(x, scale), f"rmsnorm_fwd {M}x{H}")
    dy = torch.randn(M, H, generator=gen).to(BF16)
    dres = torch.randn(M, H, generator=gen).to(BF16)
    xf = x.float().requires_grad_(True)
    sf = scale.clone().requires_grad_(True)
    ms = (xf * xf).mean(-1, keepdim=True)
    (xf * (torch.rsqrt(ms + 1e-6) * sf) * dy.float()).sum().backward()
    dscale = torch.zeros(H, dtype=torch.float32, device=DEV)
    dx = ops.rmsnorm_bwd(dy.to(DEV), x.to(DEV), scale.to(DEV), dscale=dscale, dres=dres.to(DEV))
    ref_dx = (xf.grad.to(BF16).float() + dres.float()).to(BF16)
    assert_close_bf16(dx, ref_dx, f"rmsnorm_bwd dx {M}x{H}", rtol=2e-2)
    torch.testing.assert_close(dscale.cpu(), sf.grad, rtol=2e-3, atol=2e-3 * sf.grad.abs().max().item())


def test_glu_bwd_and_gelu_bwd(ops):
    gen = g(77)
    M, F = 300, 768
    u = torch.randn(M, F, generator=gen).to(BF16)
    gt = torch.randn(M, F, generator=gen).to(BF16)
    dh = torch.randn(M, F, generator=gen).to(BF16)
    uf, gf = u.float().requires_grad_(True), gt.float().requires_grad_(True)
    c = math.sqrt(2.0 / math.pi)
    a = 0.5 * uf * (1.0 + torch.tanh(c * (uf + 0.044715 * uf ** 3)))
    (a * gf * dh.float()).sum().backward()
    dug = ops.glu_bwd(dh.to(DEV), u.to(DEV), gt.to(DEV))
    assert_close_bf16(dug[:, :F], uf.grad, "glu_bwd du", rtol=3e-2, abs_floor_frac=2e-2)
    assert_close_bf16(dug[:, F:], gf.grad, "glu_bwd dg", rtol=3e-2, abs_floor_frac=2e-2)
    z = torch.randn(M, F, generator=gen).to(BF16)
    zf = z.float().requires_grad_(True)
    (0.5 * zf * (1.0 + torch.tanh(c * (zf + 0.044715 * zf ** 3))) * dh.float()).sum().backward()
    dz = ops.gelu_bwd(dh.to(DEV), z.to(DEV))
    assert_close_bf16(dz, zf.grad, "gelu_bwd", rtol=2e-2, abs_floor_frac=1e-2)


def test_rope_inverse_is_transpose(ops):
    gen = g(5)
    M, nh, D = 200, 5, 32
    x = torch.randn(M, nh * D + 32, generator=gen).to(BF16)       # extra 32 untouched columns
    pos = torch.randint(0, 512, (M,), generator=gen).to(torch.int32)
    xd = x.clone().to(DEV)
    ops.rope_(xd, nh * D + 32, pos.to(DEV), _rope_timescale(D).to(DEV), M, nh, D, inverse=False)
    ref = O.apply_rope(x[:, :nh * D].reshape(M, 1, nh, D), pos[:, None], D).reshape(M, nh * D)
    assert_close_bf16(xd[:, :nh * D], ref, "rope fwd")
    assert torch.equal(xd[:, nh * D:].cpu(), x[:, nh * D:])
    ops.rope_(xd, nh * D + 32, pos.to(DEV), _rope_timescale(D).to(DEV), M, nh, D, inverse=True)
    assert_close_bf16(xd[:, :nh * D], x[:, :nh * D], "rope inverse(rope(x)) == x", rtol=2e-2, abs_floor_frac=1e-2)


@pytest.mark.parametrize("M,A", [(3003, 8), (5, 3), (4097, 12), (131072, 8)])
def test_policy_logits_shapes(ops, M, A):
    """Two tokens per lane group and iteration: ragged tails, the A <= 8 and generic instantiations, no mask."""
    gen = g(M + A)
    H = 128
    x = torch.randn(M, H, generator=gen).to(BF16)
    table = torch.randn(A, H, generator=gen) * 0.1
    logits = ops.policy_logits(x.to(DEV), table.to(DEV), None)
    torch.testing.assert_close(logits.cpu(), x.float() @ table.t(), rtol=1e-4, atol=1e-4)


def test_policy_logits_and_ppo_loss(ops):
    gen = g(31)
    M, H, A = 3000, 128, 6
    x = torch.randn(M, H, generator=gen).to(BF16)
    table = torch.randn(A, H, generator=gen) * 0.1
    mask = torch.rand(M, A, generator=gen) < 0.8
    mask[:, 0] = True
    logits = ops.policy_logits(x.to(DEV), table.to(DEV), mask.to(DEV))
    ref = torch.where(mask, x.float() @ table.t(), torch.tensor(torch.finfo(torch.float32).min))
    torch.testing.assert_close(logits.cpu(), ref, rtol=1e-4, atol=1e-4)
    actions = torch.zeros(M, dtype=torch.int32)
    old_lp = -torch.rand(M, generator=gen) * 2
    adv = torch.randn(M, generator=gen)
    lg = ref.clone().requires_grad_(True)
    lp = O.categorical_log_prob(lg, actions)
    ratio = torch.exp(lp - old_lp)
    actor = -torch.minimum(ratio * adv, torch.clamp(ratio, 1 - 0.28, 1 + 0.28) * adv).mean()
    ent_loss = -O.categorical_entropy(lg).mean()
    (actor + 0.001 * ent_loss).backward()
    losses, dl = ops.ppo_policy_loss(logits, actions.to(DEV), old_lp.to(DEV), adv.to(DEV), 0.28, 0.001)
    torch.testing.assert_close(losses.cpu(), torch.stack([actor.detach(), ent_loss.detach()]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dl.cpu(), lg.grad, rtol=1e-3, atol=1e-8)
    # tied-table transpose
    dtable = torch.zeros(A, H, dtype=torch.float32, device=DEV)
    dx = ops.policy_head_bwd(dl, x.to(DEV), table.to(DEV), dtable)
    assert_close_bf16(dx, dl.cpu() @ table, "policy_head_bwd dx", rtol=2e-2, abs_floor_frac=1e-2)
    torch.testing.assert_close(dtable.cpu(), dl.cpu().t() @ x.float(), rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("M,H,A", [(3001, 128, 8), (7, 128, 5), (70001, 128, 8), (2000, 256, 6), (1500, 128, 12)])
def test_policy_head_bwd_shapes(ops, M, H, A):
    """Tied-table transpose with the value path's cotangent added (dx_other): H = 128 takes the 4-consecutive-column
    path with two rows in flight (ragged M), other widths the strided path."""
    gen = g(M + H + A)
    x = torch.randn(M, H, generator=gen).to(BF16)
    table = torch.randn(A, H, generator=gen) * 0.1
    dl = torch.randn(M, A, generator=gen) * 0.01
    other = (torch.randn(M, H, generator=gen) * 0.01).to(BF16)
    dtable = torch.zeros(A, H, dtype=torch.float32, device=DEV)
    dx = ops.policy_head_bwd(dl.to(DEV), x.to(DEV), table.to(DEV), dtable, dx_other=other.to(DEV))
    ref = ((dl @ table).to(BF16).float() + other.float()).to(BF16)
    assert_close_bf16(dx, ref, f"policy_head_bwd dx M={M} H={H} A={A}", rtol=2e-2, abs_floor_frac=1e-2)
    torch.testing.assert_close(dtable.cpu(), dl.t() @ x.float(), rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("M,H,A,NT", [(5000, 128, 8, 0), (131, 128, 6, 3), (70003, 128, 8, 0), (900, 256, 8, 4), (333, 64, 5, 2)])
def test_embed_bwd(ops, M, H, A, NT):
    """Transpose of the embedding sum (network.py:303-309): reward kernel / bias and the action / task table rows, with
    wrapped negative and out-of-range indices (jnp.take mode=fill), accumulated INTO the gradient buffers.  H = 128 and
    64 take the 8-columns-per-thread kernel, H = 256 with task rows the column-per-thread one."""
    gen = g(M + H + A + NT)
    dx = (torch.randn(M, H, generator=gen) * 0.1).to(BF16)
    reward = torch.randn(M, generator=gen)
    act = torch.randint(-2, A + 2, (M,), generator=gen).to(torch.int32)
    task = torch.randint(-1, NT + 1, (M,), generator=gen).to(torch.int32) if NT else None
    d_wr = torch.full((H,), 0.5, device=DEV)
    d_br = torch.full((H,), -0.25, device=DEV)
    d_at = torch.full((A, H), 1.0, device=DEV)
    d_tt = torch.full((NT, H), 2.0, device=DEV) if NT else None
    ops.embed_bwd(dx.to(DEV), reward.to(DEV), act.to(DEV), task.to(DEV) if NT else None, A, NT, d_wr, d_br, d_at, d_tt)
    g32 = dx.double()
    sq = float(torch.tensor(math.sqrt(H)).to(BF16))
    rb = reward.to(BF16).double()
    torch.testing.assert_close(d_wr.cpu().double(), 0.5 + (rb[:, None] * g32).sum(0), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(d_br.cpu().double(), -0.25 + g32.sum(0), rtol=1e-4, atol=1e-4)

    def table_ref(idx, n, base):
        idx = idx.long()
        idx = torch.where(idx < 0, idx + n, idx)
        ref = torch.full((n, H), base, dtype=torch.float64)
        ok = (idx >= 0) & (idx < n)
        ref.index_add_(0, idx[ok], g32[ok] * sq)
        return ref
    torch.testing.assert_close(d_at.cpu().double(), table_ref(act, A, 1.0), rtol=1e-4, atol=1e-4)
    if NT:
        torch.testing.assert_close(d_tt.cpu().double(), table_ref(task, NT, 2.0), rtol=1e-4, atol=1e-4)


def test_value_head_hilo(ops):
    """fp32-weight Linear emulated with a bf16 hi/lo split (values.py:34): ~fp32 accuracy on bf16 activations."""
    gen = g(41)
    M, K, N = 500, 768, 51
    x = torch.randn(M, K, generator=gen).to(BF16)
    W = torch.randn(K, N, generator=gen) / math.sqrt(K)
    b = torch.randn(N, generator=gen) * 0.1
    hi = W.t().to(BF16)
    lo = (W.t() - hi.float()).to(BF16)
    wt = torch.zeros(128, K, dtype=BF16)
    wt[:N] = hi
    wt[64:64 + N] = lo
    out = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.linear_f32w(x.to(DEV), wt.to(DEV), b.to(DEV), out, N)
    ref = x.float() @ W + b
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-4)


def test_transpose_tb(ops):
    gen = g(3)
    T, B = 7, 12
    src = torch.randint(0, 255, (T, B, 11, 11, 2), generator=gen).to(torch.uint8)
    perm = torch.randperm(B, generator=gen).to(torch.int32)
    dst = torch.empty(B, T, 11, 11, 2, dtype=torch.uint8, device=DEV)
    ops.transpose_tb(src.to(DEV), dst, perm.to(DEV), T, B)
    assert torch.equal(dst.cpu(), src[:, perm.long()].permute(1, 0, 2, 3, 4))
    f = torch.randn(T, B, generator=gen)
    fd = torch.empty(B, T, device=DEV)
    ops.transpose_tb(f.to(DEV), fd, None, T, B)
    assert torch.equal(fd.cpu(), f.t())


@pytest.mark.parametrize("T,B,shape", [(512, 9, (11, 11, 2)), (136, 5, (11, 11, 2)), (64, 33, (5, 5, 2)), (8, 3, (11, 11, 2)),
                                       (256, 4, (7, 7, 3))])
def test_transpose_tb_row_assembler(ops, T, B, shape):
    """Even item sizes that are not 4-byte multiples (uint8 observations) go through the shared-memory row assembler
    (full and ragged 128-step segments); odd sizes keep the byte kernel.  Bit exact against torch indexing."""
    gen = g(T + B)
    src = torch.randint(0, 255, (T, B) + shape, generator=gen).to(torch.uint8)
    perm = torch.randperm(B, generator=gen).to(torch.int32)
    dst = torch.full((B, T) + shape, 7, dtype=torch.uint8, device=DEV)
    ops.transpose_tb(src.to(DEV), dst, perm.to(DEV), T, B)
    assert torch.equal(dst.cpu(), src[:, perm.long()].transpose(0, 1))
    dst.fill_(9)
    ops.transpose_tb(src.to(DEV), dst, None, T, B)
    assert torch.equal(dst.cpu(), src.transpose(0, 1))


@pytest.mark.parametrize("norm", [False, True])
@pytest.mark.parametrize("M,F,cluster", [(77, 768, 0), (4096, 768, 0), (4096, 768, 1), (4096, 768, 2), (1000, 512, 4),
                                         (5000, 768, 4), (20000, 768, 0), (20000, 640, 2), (300, 192, 0)])
def test_glu_fused(ops, M, F, cluster, norm):
    """Fused GLU block + residual (h stays on chip) vs the oracle GLUBlock; every cluster split of the hidden
    dimension (partials summed over distributed shared memory) and the persistent multi-tile loop."""
    from jaxrl_b200._lib import lib
    gen = g(M + F + cluster)
    K = 128
    x = torch.randn(M, K, generator=gen).to(BF16)
    res = torch.randn(M, K, generator=gen).to(BF16)
    wu = torch.randn(K, F, generator=gen) / math.sqrt(K)
    wg = torch.randn(K, F, generator=gen) / math.sqrt(K)
    wd = torch.randn(F, K, generator=gen) / math.sqrt(F)
    pdict = {"f.up_proj.kernel": wu, "f.up_gate.kernel": wg, "f.down_proj.kernel": wd}
    scale = 1.0 + 0.1 * torch.randn(K, generator=gen)
    xin = O.rms_norm(x, scale) if norm else x
    ref = (O.glu_block(pdict, "f.", xin).float() + res.float()).to(BF16)
    out = torch.empty(M, K, dtype=BF16, device=DEV)
    lib.mpx_debug_set(3, cluster)
    try:
        ops.glu_fused(x.to(DEV), ops.pack_glu_weight64(wu, wg).to(DEV), wd.t().contiguous().to(BF16).to(DEV), res.to(DEV), out, F,
                      norm_scale=scale.to(DEV) if norm else None)
        torch.cuda.synchronize()
    finally:
        lib.mpx_debug_set(3, 0)
    assert_close_bf16(out, ref, f"glu_fused M={M} F={F} cl={cluster}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=1e-3,
                      max_err_frac=0.1)


@pytest.mark.parametrize("M,Vh,NL,cluster,norm", [(77, 768, 51, 0, False), (4096, 768, 51, 0, True), (4096, 768, 51, 1, False),
                                                  (4096, 768, 51, 2, True), (1000, 256, 51, 4, True), (20000, 768, 51, 0, False),
                                                  (20000, 768, 51, 4, True), (300, 192, 10, 0, True), (129, 64, 64, 0, False)])
def test_value_head_fused(ops, M, Vh, NL, cluster, norm):
    """Fused value path (RMSNorm -> value MLP + gelu -> fp32 value head -> HL-Gauss expectation) vs the oracle."""
    from jaxrl_b200._lib import lib
    gen = g(M + Vh + NL + cluster)
    H = 128
    x = torch.randn(M, H, generator=gen).to(BF16)
    scale = 1.0 + 0.1 * torch.randn(H, generator=gen)
    wvm = torch.randn(H, Vh, generator=gen) / math.sqrt(H)
    bvm = torch.randn(Vh, generator=gen) * 0.1
    wvh = torch.randn(Vh, NL, generator=gen) / math.sqrt(Vh)
    bvh = torch.randn(NL, generator=gen) * 0.1
    _, centers = O.calculate_supports(-10.0, 10.0, NL)
    xf = O.rms_norm(x, scale) if norm else x
    pv = O.gelu_tanh(O.linear(xf, wvm, bvm))
    vl_ref = O.linear(pv, wvh, bvh, dtype=None)
    val_ref = O.hlgauss_get_value(vl_ref, centers)
    wt_vm = wvm.t().contiguous().to(BF16)
    hi = wvh.t().contiguous().to(BF16)
    lo = (wvh.t().contiguous() - hi.float()).to(BF16)
    hilo = torch.zeros(128, Vh, dtype=BF16)
    hilo[:NL] = hi
    hilo[64:64 + NL] = lo
    value = torch.empty(M, dtype=torch.float32, device=DEV)
    vlogits = torch.empty(M, NL, dtype=torch.float32, device=DEV)
    lib.mpx_debug_set(4, cluster)
    try:
        ops.value_head_fused(x.to(DEV), wt_vm.to(DEV), bvm.to(BF16).to(DEV), hilo.to(DEV), bvh.to(DEV), centers.to(DEV),
                             value, vlogits=vlogits, norm_scale=scale.to(DEV) if norm else None)
        torch.cuda.synchronize()
    finally:
        lib.mpx_debug_set(4, 0)
    assert_close_bf16(vlogits, vl_ref, f"value_head_fused vlogits M={M} Vh={Vh} cl={cluster} norm={norm}", rtol=1e-2,
                      abs_floor_frac=2e-2, allow_frac=2e-3, max_err_frac=0.1)
    torch.testing.assert_close(value.cpu(), val_ref, rtol=1e-2, atol=3e-2)
    # internal consistency: value == get_value(vlogits) to fp32 accuracy
    torch.testing.assert_close(value.cpu(), O.hlgauss_get_value(vlogits.cpu(), centers), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("M,H,A,norm", [(4096, 128, 6, True), (777, 128, 8, False), (3000, 256, 19, True), (5, 64, 1, True)])
def test_policy_head_sample_fused(ops, M, H, A, norm):
    """Fused (RMSNorm ->) logits -> mask -> Gumbel arg-max -> log-prob vs the separate kernels and the oracle."""
    gen = g(M + H + A)
    x = torch.randn(M, H, generator=gen).to(BF16)
    scale = 1.0 + 0.1 * torch.randn(H, generator=gen)
    table = torch.randn(A, H, generator=gen) * 0.1
    mask = torch.rand(M, A, generator=gen) < 0.8
    mask[:, 0] = True
    xd, td, md = x.to(DEV), table.to(DEV), mask.to(DEV)
    xf_ref = O.rms_norm(x, scale) if norm else x
    ref_logits = torch.where(mask, xf_ref.float() @ table.t(), torch.tensor(torch.finfo(torch.float32).min))
    # (1) caller-provided noise: same arg-max as the oracle wherever the top-2 margin is not a rounding tie
    gum = -torch.log(-torch.log(torch.rand(M, A, generator=gen).clamp(1e-6, 1 - 1e-6)))
    act = torch.empty(M, dtype=torch.int32, device=DEV)
    lp = torch.empty(M, dtype=torch.float32, device=DEV)
    logits = torch.empty(M, A, dtype=torch.float32, device=DEV)
    xf = torch.empty(M, H, dtype=BF16, device=DEV)
    ops.policy_head_sample(xd, td, act, lp, mask=md, gumbel_in=gum.to(DEV), norm_scale=scale.to(DEV) if norm else None,
                           logits_out=logits, xf_out=xf if norm else None)
    torch.cuda.synchronize()
    live = mask
    torch.testing.assert_close(logits.cpu()[live], ref_logits[live], rtol=1e-2, atol=2e-2)
    assert (logits.cpu()[~live] == torch.finfo(torch.float32).min).all()
    if norm:
        assert_close_bf16(xf, xf_ref, "policy_head_sample xf")
    # exact self-consistency: the action is the arg-max of the kernel's own logits + the given noise
    own = O.categorical_sample_gumbel(logits.cpu(), gum)
    assert torch.equal(act.cpu(), own)
    torch.testing.assert_close(lp.cpu(), O.categorical_log_prob(logits.cpu(), own), rtol=1e-5, atol=1e-5)
    # (2) built-in Philox noise == mpx_gumbel_noise with the same (seed, stream) + mpx_policy_sample
    off = torch.tensor([3], dtype=torch.int64, device=DEV)
    act2 = torch.empty_like(act)
    lp2 = torch.empty_like(lp)
    ops.policy_head_sample(xd, td, act2, lp2, seed=1234, stream_id=7, stream_offset=off, mask=md,
                           norm_scale=scale.to(DEV) if norm else None)
    noise = ops.gumbel_noise(torch.empty(M, A, dtype=torch.float32, device=DEV), 1234, 7, off)
    act3, lp3 = ops.policy_sample(logits, noise)
    assert torch.equal(act2, act3)
    torch.testing.assert_close(lp2, lp3, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("M,sizes,tasks", [(32, (5, 5), 0), (4096, (5, 5), 0), (77, (3, 2, 2), 3), (9000, (5, 5), 0), (50, (6,), 0)])
def test_obs_embed_fused(ops, M, sizes, tasks):
    """Fused observation encoder (one-hot conv gather + two tcgen05 convolutions on no-swizzle operands) + embedding sum
    vs the oracle obs_encoder / model_embed."""
    from jaxrl_b200.config import ModelConfig
    gen = g(M + len(sizes) + tasks)
    K, C, H, A = len(sizes), sum(sizes), 128, 6
    cfg = ModelConfig(obs_shape=(11, 11, K), obs_max_value=tuple(sizes), action_dim=A, task_count=max(tasks, 1))
    obs = torch.stack([torch.randint(0, n, (M, 11, 11), generator=gen) for n in sizes], -1).to(torch.uint8)
    p = {"obs_encoder.layers.0.kernel": torch.randn(3, 3, C, 16, generator=gen) * 0.3,
         "obs_encoder.layers.0.bias": torch.randn(16, generator=gen) * 0.1,
         "obs_encoder.layers.1.kernel": torch.randn(3, 3, 16, 32, generator=gen) / 12.0,
         "obs_encoder.layers.1.bias": torch.randn(32, generator=gen) * 0.1,
         "obs_encoder.layers.2.kernel": torch.randn(3, 3, 32, H, generator=gen) / 17.0,
         "obs_encoder.layers.2.bias": torch.randn(H, generator=gen) * 0.1,
         "reward_encoder.kernel": torch.randn(1, H, generator=gen), "reward_encoder.bias": torch.randn(H, generator=gen) * 0.1,
         "action_embedder.embedding_table": torch.randn(A, H, generator=gen) * 0.1}
    if tasks:
        p["task_embedder.embedding_table"] = torch.randn(tasks, H, generator=gen) * 0.1
    reward = torch.randn(M, generator=gen)
    la = torch.randint(-1, A, (M,), generator=gen).to(torch.int32)
    tk = torch.randint(0, tasks, (M,), generator=gen).to(torch.int32) if tasks else None
    emb_ref = O.obs_encoder(p, cfg, obs)
    ts = O.TimeStep(obs=obs[:, None], time=torch.zeros(M, 1, dtype=torch.int32), terminated=torch.zeros(M, 1, dtype=torch.bool),
                    last_action=la[:, None], reward=reward[:, None], action_mask=None, task_ids=None if tk is None else tk[:, None])
    x_ref = O.model_embed(p, cfg, ts)[:, 0]
    d = lambda t: t.to(DEV)
    b = lambda t: t.to(BF16).to(DEV)
    w1 = b(p["obs_encoder.layers.0.kernel"].reshape(9 * C, 16))
    wt2 = b(p["obs_encoder.layers.1.kernel"].reshape(144, 32).t().contiguous())
    wt3 = b(p["obs_encoder.layers.2.kernel"].reshape(288, H).t().contiguous())
    out = torch.empty(M, H, dtype=BF16, device=DEV)
    emb = torch.empty(M, H, dtype=BF16, device=DEV)
    import os
    from jaxrl_b200._lib import lib
    image = ops.obs_fused_prep(sizes, w1, b(p["obs_encoder.layers.0.bias"]), wt2, b(p["obs_encoder.layers.1.bias"]), wt3,
                               b(p["obs_encoder.layers.2.bias"]), ops.obs_fused_image(DEV))
    ops.obs_embed_fused(d(obs), sizes, image, d(reward), d(la), None if tk is None else d(tk),
                        d(p["reward_encoder.kernel"].reshape(-1)), d(p["reward_encoder.bias"]),
                        d(p["action_embedder.embedding_table"]), d(p["task_embedder.embedding_table"]) if tasks else None,
                        out, M, 11, 11, obs_emb=emb)
    torch.cuda.synchronize()
    assert_close_bf16(emb, emb_ref, f"obs_embed_fused encoder output M={M}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=2e-3,
                      max_err_frac=0.1)
    assert_close_bf16(out, x_ref, f"obs_embed_fused x M={M}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=2e-3, max_err_frac=0.1)


@pytest.mark.parametrize("M,sizes", [(37, (5, 5)), (3000, (5, 5)), (500, (3, 2, 2)), (64, (6,))])
def test_conv1_onehot_wgrad(ops, M, sizes):
    """dW1 += onehot_im2col(obs)^T @ dz with the one-hot operand generated in shared memory, vs the materialised
    one-hot GEMM in fp64 (bf16 inputs, fp32 accumulation: exact up to summation order)."""
    gen = g(M + len(sizes))
    IH = IW = 11
    C, cout = sum(sizes), 16
    obs = torch.stack([torch.randint(0, n, (M, IH, IW), generator=gen) for n in sizes], -1).to(torch.uint8)
    dz = (torch.randn(M * 25, cout, generator=gen)).to(BF16)
    x = O.concat_one_hot(obs, sizes).float()                                  # (M, 11, 11, C)
    cols = torch.nn.functional.unfold(x.permute(0, 3, 1, 2), kernel_size=3, stride=2)   # (M, C*9, 25), index c*9 + tap
    cols = cols.reshape(M, C, 9, 25).permute(0, 3, 2, 1).reshape(M * 25, 9 * C)         # rows (m, q), features (tap, c)
    ref = cols.double().t() @ dz.double()
    dw = torch.full((9 * C, cout), 0.5, dtype=torch.float32, device=DEV)    # accumulates into existing values
    ops.conv1_onehot_wgrad(obs.to(DEV), dz.to(DEV), dw, M, IH, IW, sizes, 3, 3, 2, 2, cout)
    torch.cuda.synchronize()
    torch.testing.assert_close(dw.cpu().double() - 0.5, ref, rtol=1e-4, atol=1e-3 * max(1.0, math.sqrt(M)))


@pytest.mark.parametrize("M,F,cluster", [(77, 768, 0), (4096, 768, 0), (4096, 768, 1), (4096, 768, 2), (20000, 768, 0),
                                         (20000, 512, 4), (300, 192, 0)])
def test_glu_fused_with_out_projection(ops, M, F, cluster):
    """Attention output projection + residual + ffn_norm + GLU + residual in one kernel vs the oracle block tail."""
    from jaxrl_b200._lib import lib
    gen = g(M + F + cluster + 7)
    H = 128
    ao = torch.randn(M, H, generator=gen).to(BF16)
    x = torch.randn(M, H, generator=gen).to(BF16)
    wo = torch.randn(H, H, generator=gen) / math.sqrt(H)            # (in = heads*dim, out = H)
    wu = torch.randn(H, F, generator=gen) / math.sqrt(H)
    wg = torch.randn(H, F, generator=gen) / math.sqrt(H)
    wd = torch.randn(F, H, generator=gen) / math.sqrt(F)
    scale = 1.0 + 0.1 * torch.randn(H, generator=gen)
    x1 = (O.linear(ao, wo).float() + x.float()).to(BF16)
    pdict = {"f.up_proj.kernel": wu, "f.up_gate.kernel": wg, "f.down_proj.kernel": wd}
    ref = (O.glu_block(pdict, "f.", O.rms_norm(x1, scale)).float() + x1.float()).to(BF16)
    out = torch.empty(M, H, dtype=BF16, device=DEV)
    lib.mpx_debug_set(3, cluster)
    try:
        ops.glu_fused(ao.to(DEV), ops.pack_glu_weight64(wu, wg).to(DEV), wd.t().contiguous().to(BF16).to(DEV), x.to(DEV), out, F,
                      norm_scale=scale.to(DEV), wt_o=wo.t().contiguous().to(BF16).to(DEV))
        torch.cuda.synchronize()
    finally:
        lib.mpx_debug_set(3, 0)
    assert_close_bf16(out, ref, f"glu_fused+proj M={M} F={F} cl={cluster}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=2e-3,
                      max_err_frac=0.1)


@pytest.mark.parametrize("M,F,A,masked,philox", [(4096, 768, 6, True, True), (77, 768, 8, False, False), (4000, 512, 3, True, False),
                                                 (20000, 768, 6, True, True), (300, 192, 19, False, True), (1, 256, 2, True, True)])
def test_glu_fused_head_matches_separate_kernels(ops, M, F, A, masked, philox):
    """mpx_glu_fused_head (last feed-forward half + policy head in one launch; by-rows cluster reduction) is bit-identical to
    mpx_glu_fused followed by mpx_policy_head_sample: block output, normalised rows, logits, sampled action, log-prob.
    Covers the fused epilogue (cluster of 4, A <= 8) and the shapes where the library issues the two kernels itself
    (M = 20000: no cluster of 4; A = 19)."""
    gen = g(M + F + A)
    H = 128
    ao = torch.randn(M, H, generator=gen).to(BF16).to(DEV)
    x = torch.randn(M, H, generator=gen).to(BF16).to(DEV)
    wo = (torch.randn(H, H, generator=gen) / math.sqrt(H)).t().contiguous().to(BF16).to(DEV)
    wu = torch.randn(H, F, generator=gen) / math.sqrt(H)
    wg = torch.randn(H, F, generator=gen) / math.sqrt(H)
    wug = ops.pack_glu_weight64(wu, wg).to(DEV)
    wd = (torch.randn(F, H, generator=gen) / math.sqrt(F)).t().contiguous().to(BF16).to(DEV)
    scale = (1.0 + 0.1 * torch.randn(H, generator=gen)).to(DEV)
    oscale = (1.0 + 0.1 * torch.randn(H, generator=gen)).to(DEV)
    table = (torch.randn(A, H, generator=gen) * 0.1).to(DEV)
    mask = None
    if masked:
        mask = torch.rand(M, A, generator=gen) < 0.8
        mask[:, 0] = True
        mask = mask.to(DEV)
    gum = None if philox else (-torch.log(-torch.log(torch.rand(M, A, generator=gen).clamp(1e-6, 1 - 1e-6)))).to(DEV)
    off = torch.tensor([5], dtype=torch.int64, device=DEV) if philox else None

    def outs():
        return dict(out=torch.full((M + 2, H), 7.0, dtype=BF16, device=DEV), xf=torch.empty(M, H, dtype=BF16, device=DEV),
                    logits=torch.empty(M, A, dtype=torch.float32, device=DEV), act=torch.full((M + 2,), -9, dtype=torch.int32, device=DEV),
                    lp=torch.full((M + 2,), -9.0, dtype=torch.float32, device=DEV))
    a, b = outs(), outs()
    ops.glu_fused(ao, wug, wd, x, a["out"][:M], F, norm_scale=scale, wt_o=wo)
    ops.policy_head_sample(a["out"][:M], table, a["act"][:M], a["lp"][:M], seed=99, stream_id=11, stream_offset=off, mask=mask,
                           gumbel_in=gum, norm_scale=oscale, logits_out=a["logits"], xf_out=a["xf"])
    ops.glu_fused_head(ao, wug, wd, x, b["out"][:M], F, scale, wo, oscale, table, b["act"][:M], b["lp"][:M], seed=99, stream_id=11,
                       stream_offset=off, mask=mask, gumbel_in=gum, logits_out=b["logits"], xf_out=b["xf"])
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), f"glu_fused_head differs from the two-kernel path in {k} (M={M} F={F} A={A})"
    assert (b["act"][M:] == -9).all() and (b["out"][M:] == 7.0).all()        # nothing written past row M
    assert ((b["act"][:M] >= 0) & (b["act"][:M] < A)).all()


@pytest.mark.parametrize("M,F", [(128, 128), (77, 256), (1000, 768), (20000, 768), (131072, 768), (300, 384)])
def test_glu_bwd_fused(ops, M, F):
    """One-kernel GLU backward (u, g, dh recomputed in TMEM) vs the unfused kernel chain (glu_up(save) -> linear ->
    glu_bwd -> linear), which test_glu_up / test_glu_bwd pin to the oracle, and vs fp32 autograd of the oracle block."""
    gen = g(M + F + 3)
    H = 128
    xn = torch.randn(M, H, generator=gen).to(BF16)
    dy = (torch.randn(M, H, generator=gen) * 0.5).to(BF16)
    wu = torch.randn(H, F, generator=gen) / math.sqrt(H)
    wg = torch.randn(H, F, generator=gen) / math.sqrt(H)
    wd = torch.randn(F, H, generator=gen) / math.sqrt(F)
    wt_ug64 = ops.pack_glu_weight64(wu, wg).to(DEV)
    wt_d = wd.t().contiguous().to(BF16).to(DEV)
    dug, dxn = ops.glu_bwd_fused(xn.to(DEV), dy.to(DEV), wt_ug64, wt_d, F)
    # the training forward: same kernel as the rollout's, plus the h output kept for dWdown
    res = torch.randn(M, H, generator=gen).to(BF16)
    y = torch.empty(M, H, dtype=BF16, device=DEV)
    h = torch.full((M, F), float("nan"), dtype=BF16, device=DEV)
    ops.glu_fused(xn.to(DEV), wt_ug64, wt_d, res.to(DEV), y, F, h_out=h)
    torch.cuda.synchronize()
    # unfused chain on the same device
    h_ref, u, gg = ops.glu_up(xn.to(DEV), ops.pack_glu_weight(wu, wg).to(DEV), F, save=True)
    dh = ops.linear(dy.to(DEV), wd.to(BF16).to(DEV))                  # (M,H) @ (F,H)^T
    dug_ref = torch.empty(M, 2 * F, dtype=BF16, device=DEV)
    ops.glu_bwd(dh, u, gg, out=dug_ref)
    w_ug = torch.cat([wu, wg], dim=1).to(BF16).to(DEV)                # (H, 2F)
    dxn_ref = ops.linear(dug_ref, w_ug)
    y_ref = ops.linear(h_ref, wt_d, residual=res.to(DEV))
    torch.cuda.synchronize()
    assert_close_bf16(y, y_ref.cpu(), f"glu_fused(h_out) y M={M}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=1e-3, max_err_frac=0.1)
    assert_close_bf16(h, h_ref.cpu(), f"glu_bwd_fused h M={M}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=1e-3, max_err_frac=0.1)
    assert_close_bf16(dug, dug_ref.cpu(), f"glu_bwd_fused dug M={M}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=1e-3, max_err_frac=0.1)
    assert_close_bf16(dxn, dxn_ref.cpu(), f"glu_bwd_fused dxn M={M}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=1e-3, max_err_frac=0.1)
    if M <= 1000:
        # fp32 autograd of the oracle block: loose, it does not model the bf16 roundings of the backward intermediates
        xf = xn.float().requires_grad_(True)
        pdict = {"f.up_proj.kernel": wu.to(BF16).float(), "f.up_gate.kernel": wg.to(BF16).float(), "f.down_proj.kernel": wd.to(BF16).float()}
        up = xf @ pdict["f.up_proj.kernel"]
        gate = xf @ pdict["f.up_gate.kernel"]
        y = (O.gelu_tanh(up) * gate) @ pdict["f.down_proj.kernel"]
        (y * dy.float()).sum().backward()
        err = (dxn.float().cpu() - xf.grad).abs().max().item()
        assert err < 0.05 * xf.grad.abs().max().item() + 0.05, err


@pytest.mark.parametrize("geom", [(5, 5, 16, 3, 3, 1, 1), (11, 11, 8, 3, 3, 2, 2), (6, 7, 8, 2, 3, 1, 2)])
@pytest.mark.parametrize("with_z", [False, True])
def test_col2im(ops, geom, with_z):
    """Transpose of im2col (scatter-sum of patches), optionally fused with the gelu transpose of the layer below."""
    IH, IW, C, kh, kw, sh, sw = geom
    gen = g(sum(geom) + with_z)
    M = 300
    OH, OW = (IH - kh) // sh + 1, (IW - kw) // sw + 1
    dx = torch.randn(M * OH * OW, kh * kw * C, generator=gen).to(BF16)
    z = torch.randn(M, IH, IW, C, generator=gen).to(BF16)
    # reference: fold (fp32 sum) -> bf16 -> * gelu'(z) -> bf16
    cols = dx.float().reshape(M, OH * OW, kh * kw, C).permute(0, 3, 2, 1).reshape(M, C * kh * kw, OH * OW)
    ref = torch.nn.functional.fold(cols, (IH, IW), (kh, kw), stride=(sh, sw)).permute(0, 2, 3, 1).to(BF16)
    if with_z:
        zf = z.float().requires_grad_(True)
        O.gelu_tanh(zf).sum().backward()
        ref = (ref.float() * zf.grad).to(BF16)
    out = torch.empty(M, IH, IW, C, dtype=BF16, device=DEV)
    ops.col2im(dx.to(DEV), out, M, IH, IW, C, kh, kw, sh, sw, z=z.to(DEV) if with_z else None)
    assert_close_bf16(out, ref, f"col2im {geom} z={with_z}", rtol=1e-2, abs_floor_frac=4e-3)
