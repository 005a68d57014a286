"""GPU: size-independent properties at BASELINE.json's full sizes (B = 4096 agents, T = 512, 131 072-token minibatches),
where the CPU oracle is too slow to be the checker: linearity / closed forms of the GAE scan, normalisation and zero-sum
identities of the HL-Gauss loss, convexity / constant-value identities of the attention kernels, causality."""

import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import rope_timescale  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16
B, T = 4096, 512


@pytest.fixture(scope="module")
def ops():
    from jaxrl_b200 import ops as _ops
    return _ops


def test_gae_full_size_closed_forms(ops):
    g = torch.Generator(device=DEV).manual_seed(1)
    r = torch.randn(B, T, device=DEV, generator=g)
    v = torch.randn(B, T + 1, device=DEV, generator=g)
    never = torch.zeros(B, T, dtype=torch.bool, device=DEV)
    # discount 0: targets == rewards exactly (reference tests/test_advantage.py:150-163)
    _, tgt, _ = ops.gae(r, v, never, 0.0, 0.95)
    assert torch.equal(tgt, r)
    # lambda = 1, no termination: targets are the discounted returns bootstrapped with V_T (closed form, fp64)
    gamma = 0.96609
    _, tgt, _ = ops.gae(r, v, never, gamma, 1.0)
    pw = gamma ** torch.arange(T + 1, device=DEV, dtype=torch.float64)
    disc = (r.double() * pw[:T]).flip(1).cumsum(1).flip(1) / pw[:T] + (gamma ** (T - torch.arange(T, device=DEV))).double() * v[:, T:].double()
    torch.testing.assert_close(tgt.double(), disc, rtol=2e-4, atol=2e-4)
    # termination everywhere: targets == rewards, advantages == rewards - values
    always = torch.ones(B, T, dtype=torch.bool, device=DEV)
    adv, tgt, _ = ops.gae(r, v, always, gamma, 0.91237)
    assert torch.equal(tgt, r) and torch.equal(adv, r - v[:, :T])


def test_gae_full_size_linearity_and_normalisation(ops):
    g = torch.Generator(device=DEV).manual_seed(2)
    term = torch.rand(B, T, device=DEV, generator=g) < (1.0 / 64)
    r1, r2 = torch.randn(B, T, device=DEV, generator=g), torch.randn(B, T, device=DEV, generator=g)
    v1, v2 = torch.randn(B, T + 1, device=DEV, generator=g), torch.randn(B, T + 1, device=DEV, generator=g)
    a1, t1, _ = ops.gae(r1, v1, term, 0.96609, 0.91237)
    a2, t2, _ = ops.gae(r2, v2, term, 0.96609, 0.91237)
    a12, t12, stats = ops.gae(2 * r1 - 3 * r2, 2 * v1 - 3 * v2, term, 0.96609, 0.91237)
    torch.testing.assert_close(a12, 2 * a1 - 3 * a2, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(t12, 2 * t1 - 3 * t2, rtol=1e-4, atol=1e-4)
    ops.adv_normalize(a12, stats)
    assert abs(a12.double().mean().item()) < 1e-5 and abs(a12.double().std(unbiased=False).item() - 1.0) < 1e-4


def test_hlgauss_full_size_identities(ops):
    n, NL = 256 * T, 51
    g = torch.Generator(device=DEV).manual_seed(3)
    logits = torch.randn(n, NL, device=DEV, generator=g) * 2
    targets = torch.randn(n, device=DEV, generator=g) * 5
    sup = torch.linspace(-10.0, 10.0, NL + 1, device=DEV)
    loss, dl = ops.hlgauss_loss(logits, targets, sup, -10.0, 10.0, 0.1038127, 1.0)
    # d loss / d logits = (softmax - p) / n : every row sums to zero, |entries| <= 1/n
    assert dl.double().sum(-1).abs().max().item() < 1e-9
    assert dl.abs().max().item() <= 1.0 / n + 1e-12
    # cross entropy >= entropy of the target distribution >= 0, and the uniform-logit loss is log(NL)
    assert loss.item() > 0
    lu, _ = ops.hlgauss_loss(torch.zeros(n, NL, device=DEV), targets, sup, -10.0, 10.0, 0.1038127, 1.0, want_grad=False)
    assert abs(lu.item() - math.log(NL)) < 1e-4
    # expectation of a one-hot-like distribution is its bin centre; uniform logits give the support's midpoint
    centers = (sup[:-1] + sup[1:]) / 2
    peaked = torch.full((n, NL), -80.0, device=DEV)
    idx = torch.randint(0, NL, (n,), device=DEV, generator=g)
    peaked[torch.arange(n, device=DEV), idx] = 80.0
    torch.testing.assert_close(ops.hlgauss_value(peaked, centers), centers[idx], rtol=1e-6, atol=1e-6)
    assert ops.hlgauss_value(torch.zeros(n, NL, device=DEV), centers).abs().max().item() < 1e-5


@pytest.mark.parametrize("step", [0, 1, 255, 511])
def test_decode_attention_full_size_identities(ops, step):
    S, Nq, Nkv, D = T, 4, 1, 32
    g = torch.Generator(device=DEV).manual_seed(4 + step)
    q = torch.randn(B, Nq * D, device=DEV, generator=g).to(BF16)
    kc = torch.randn(B, S, Nkv, D, device=DEV, generator=g).to(BF16)
    pos = torch.full((B,), step, dtype=torch.int32, device=DEV)
    # (1) every value row of an agent identical -> the output IS that row, whatever the scores (softmax sums to one)
    vrow = torch.randn(B, 1, Nkv, D, device=DEV, generator=g).to(BF16)
    vc = vrow.expand(B, S, Nkv, D).contiguous()
    out = ops.attn_decode(q, kc, vc, pos, Nq, Nkv, D).reshape(B, Nq, D).float()
    torch.testing.assert_close(out, vrow[:, 0].float().expand(B, Nq, D), rtol=2e-2, atol=2e-2)
    # (2) slots >= kv_len are never read: poisoning them with NaN changes nothing
    vc2 = torch.randn(B, S, Nkv, D, device=DEV, generator=g).to(BF16)
    ref = ops.attn_decode(q, kc, vc2, pos, Nq, Nkv, D).clone()
    if step + 1 < S:
        kc[:, step + 1:] = float("nan")
        vc2[:, step + 1:] = float("nan")
    got = ops.attn_decode(q, kc, vc2, pos, Nq, Nkv, D)
    assert torch.isfinite(got).all() and torch.equal(got, ref)
    # (3) outputs are convex combinations of the visible value rows
    lo = vc2[:, :step + 1].float().amin(1)      # (B, Nkv, D)
    hi = vc2[:, :step + 1].float().amax(1)
    o = got.reshape(B, Nq, D).float()
    assert (o >= lo - 2e-2 * lo.abs() - 2e-2).all() and (o <= hi + 2e-2 * hi.abs() + 2e-2).all()


def test_training_attention_full_size_causality(ops):
    Bm, Nq, Nkv, D = 256, 4, 1, 32
    QD, KD = Nq * D, Nkv * D
    ld = QD + 2 * KD
    g = torch.Generator(device=DEV).manual_seed(9)
    qkv = torch.randn(Bm * T, ld, device=DEV, generator=g).to(BF16)
    o1, lse1 = ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], Bm, T, Nq, Nkv, D, ld, ld, ld)
    o1, lse1 = o1.clone(), lse1.clone()
    # perturb k and v of the second half of every trajectory: outputs of the first half must not move (bit exact)
    view = qkv.view(Bm, T, ld)
    view[:, T // 2:, QD:] = torch.randn(Bm, T // 2, 2 * KD, device=DEV, generator=g).to(BF16)
    o2, lse2 = ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], Bm, T, Nq, Nkv, D, ld, ld, ld)
    a, b_ = o1.view(Bm, T, QD), o2.view(Bm, T, QD)
    assert torch.equal(a[:, :T // 2], b_[:, :T // 2]) and not torch.equal(a[:, T // 2:], b_[:, T // 2:])
    assert torch.equal(lse1.view(Bm, T, Nq)[:, :T // 2], lse2.view(Bm, T, Nq)[:, :T // 2])
    # token 0 attends to itself only: its output is its own value row
    v0 = view[:, 0, QD + KD:].float()
    torch.testing.assert_close(b_[:, 0].view(Bm, Nq, D).float(), v0[:, None].expand(Bm, Nq, D), rtol=1e-2, atol=1e-2)
    # backward: the gradient w.r.t. v summed over positions equals the dO summed over heads and positions (sum_t P = 1)
    dout = torch.randn(Bm * T, QD, device=DEV, generator=g).to(BF16)
    dqkv = torch.zeros(Bm * T, ld, dtype=BF16, device=DEV)
    pos = torch.arange(T, dtype=torch.int32, device=DEV)
    ops.attn_train_bwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], o2, dout, lse2, pos, rope_timescale(D).to(DEV), Bm, T, Nq, Nkv, D,
                       ld, ld, ld, dqkv, dqkv[:, QD:], dqkv[:, QD + KD:], ld, ld, ld)
    dv_sum = dqkv[:, QD + KD:].float().view(Bm, T, D).sum(1)
    do_sum = dout.float().view(Bm, T, Nq, D).sum((1, 2))
    torch.testing.assert_close(dv_sum, do_sum, rtol=3e-2, atol=0.5)
    assert torch.isfinite(dqkv.float()).all()
