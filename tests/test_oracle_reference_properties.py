"""The oracle pinned against every known answer / property the reference's own tests hold for the hot path
(reference tests/test_attention.py, test_rope.py, test_advantage.py, test_values.py), re-expressed for the
torch restatement.  fp32 here, as in the reference tests."""

import math

import numpy as np
import pytest
import torch

from jaxrl_b200.config import ModelConfig
from oracle import oracle as O

F32 = torch.float32


def make_attention(d_model=32, head_dim=8, num_heads=4, num_kv_heads=4, max_seq_length=16, seed=0):
    cfg = ModelConfig(hidden_features=d_model, head_dim=head_dim, num_heads=num_heads, num_kv_heads=num_kv_heads,
                      max_seq_length=max_seq_length, num_layers=1)
    g = torch.Generator().manual_seed(seed)
    p = {
        "a.query_proj.kernel": torch.randn(d_model, num_heads, head_dim, generator=g) * 0.1,
        "a.key_proj.kernel": torch.randn(d_model, num_kv_heads, head_dim, generator=g) * 0.1,
        "a.value_proj.kernel": torch.randn(d_model, num_kv_heads, head_dim, generator=g) * 0.1,
        "a.out.kernel": torch.randn(num_heads, head_dim, d_model, generator=g) * 0.1,
    }
    return cfg, p


def attn(cfg, p, x, pos, cache=None):
    return O.attention_block(p, "a.", cfg, x, pos, cache, dtype=F32)


class TestAttentionBlock:  # reference tests/test_attention.py:35-80
    def test_output_shape(self):
        cfg, p = make_attention()
        out, _ = attn(cfg, p, torch.ones(2, 8, 32), torch.arange(8)[None].repeat(2, 1))
        assert out.shape == (2, 8, 32)

    def test_causal_masking(self):
        cfg, p = make_attention()
        x = torch.randn(1, 8, 32, generator=torch.Generator().manual_seed(0))
        pos = torch.arange(8)[None]
        out1, _ = attn(cfg, p, x, pos)
        x2 = x.clone()
        x2[0, 7] = 99.0
        out2, _ = attn(cfg, p, x2, pos)
        assert torch.allclose(out1[0, :7], out2[0, :7], atol=1e-5)
        assert not torch.allclose(out1[0, 7], out2[0, 7], atol=1e-3)

    @pytest.mark.parametrize("kv", [2, 1])
    def test_gqa_mqa_shapes(self, kv):
        cfg, p = make_attention(num_heads=4, num_kv_heads=kv)
        out, _ = attn(cfg, p, torch.ones(2, 4, 32), torch.arange(4)[None].repeat(2, 1))
        assert out.shape == (2, 4, 32)


class TestKVCache:  # reference tests/test_attention.py:83-136
    def test_initialize_carry_shape(self):
        cfg, _ = make_attention(max_seq_length=16, num_kv_heads=2, head_dim=8)
        c = O.initialize_carry(cfg, 3, dtype=F32)[0]
        assert c.key.shape == (3, 16, 2, 8) and c.value.shape == (3, 16, 2, 8)

    @pytest.mark.parametrize("S,pos,slot", [(8, 3, 3), (4, 6, 2)])
    def test_cache_slot_and_wrap(self, S, pos, slot):
        cfg, p = make_attention(max_seq_length=S)
        cache = O.initialize_carry(cfg, 1, dtype=F32)[0]
        x = torch.randn(1, 1, 32, generator=torch.Generator().manual_seed(3))
        _, new = attn(cfg, p, x, torch.tensor([[pos]]), cache)
        assert new.key[0, slot].abs().sum() > 0
        other = [i for i in range(S) if i != slot]
        assert torch.all(new.key[0, other] == 0) and torch.all(new.value[0, other] == 0)

    def test_inference_matches_training_first_step(self):
        cfg, p = make_attention(d_model=32, max_seq_length=8)
        x = torch.randn(1, 1, 32, generator=torch.Generator().manual_seed(1))
        pos = torch.tensor([[0]])
        out_train, _ = attn(cfg, p, x, pos, None)
        out_inf, _ = attn(cfg, p, x, pos, O.initialize_carry(cfg, 1, dtype=F32)[0])
        assert torch.allclose(out_train, out_inf, atol=1e-5)

    def test_decode_sequence_matches_training(self):
        """Stronger than the reference test: T decode steps == causal training pass (fp32)."""
        cfg, p = make_attention(d_model=32, max_seq_length=8, num_kv_heads=2)
        x = torch.randn(2, 8, 32, generator=torch.Generator().manual_seed(5))
        pos = torch.arange(8)[None].repeat(2, 1)
        out_train, _ = attn(cfg, p, x, pos)
        cache = O.initialize_carry(cfg, 2, dtype=F32)[0]
        for t in range(8):
            o, cache = attn(cfg, p, x[:, t:t + 1], pos[:, t:t + 1], cache)
            assert torch.allclose(o[:, 0], out_train[:, t], atol=1e-5)


class TestRope:  # reference tests/test_rope.py
    def test_shape_dtype(self):
        x = torch.randn(2, 8, 4, 16)
        pos = torch.arange(8)[None].repeat(2, 1)
        assert O.apply_rope(x, pos, 16).shape == x.shape
        for dt in (torch.float32, torch.bfloat16):
            assert O.apply_rope(torch.ones(1, 4, 2, 8, dtype=dt), torch.arange(4)[None], 8).dtype == dt

    def test_preserves_norm(self):
        x = torch.randn(2, 8, 4, 16, generator=torch.Generator().manual_seed(1))
        out = O.apply_rope(x, torch.arange(8)[None].repeat(2, 1), 16)
        assert torch.allclose(x.norm(dim=-1), out.norm(dim=-1), atol=1e-5)

    def test_position_zero_identity(self):
        x = torch.randn(1, 1, 2, 8, generator=torch.Generator().manual_seed(2))
        assert torch.allclose(O.apply_rope(x, torch.zeros(1, 1, dtype=torch.int32), 8), x, atol=1e-6)

    def test_positions_differ(self):
        out = O.apply_rope(torch.ones(1, 2, 1, 8), torch.tensor([[0, 5]]), 8)
        assert not torch.allclose(out[0, 0], out[0, 1], atol=1e-3)

    def test_relative_position(self):
        q = torch.randn(1, 1, 1, 16, generator=torch.Generator().manual_seed(3))
        k = torch.randn(1, 1, 1, 16, generator=torch.Generator().manual_seed(4))
        d1 = (O.apply_rope(q, torch.tensor([[2]]), 16) * O.apply_rope(k, torch.tensor([[5]]), 16)).sum()
        d2 = (O.apply_rope(q, torch.tensor([[10]]), 16) * O.apply_rope(k, torch.tensor([[13]]), 16)).sum()
        assert torch.allclose(d1, d2, atol=1e-4)


def manual_gae(rewards, values, terminated, discount, lam):  # reference tests/test_advantage.py:21-35 (python floats)
    T = len(rewards)
    adv, tgt = [0.0] * T, [0.0] * T
    acc = values[T]
    for t in reversed(range(T)):
        d = 0.0 if terminated[t] else discount
        acc = rewards[t] + d * ((1 - lam) * values[t + 1] + lam * acc)
        tgt[t] = acc
        adv[t] = tgt[t] - values[t]
    return adv, tgt


class TestAdvantage:  # reference tests/test_advantage.py
    def test_single_step(self):
        adv, tgt = O.calculate_advantage([[1.0]], [[0.5, 0.8]], [[False]], 0.99, 0.95, False)
        assert abs(tgt[0, 0] - (1.0 + 0.99 * 0.8)) < 1e-5 and abs(adv[0, 0] - (1.0 + 0.99 * 0.8 - 0.5)) < 1e-5

    def test_termination(self):
        r, v, te = [[1.0, 2.0, 3.0]], [[0.5] * 4], [[False, True, False]]
        _, tgt = O.calculate_advantage(r, v, te, 0.99, 0.95, False)
        _, ref = manual_gae(r[0], v[0], te[0], 0.99, 0.95)
        assert np.allclose(tgt[0], ref, atol=1e-5) and abs(tgt[0, 1] - 2.0) < 1e-6

    def test_matches_loop(self):
        r, v = [[1.0, -0.5, 2.0, 0.0, 1.5]], [[0.3, 0.7, 0.1, 0.9, 0.4, 0.6]]
        adv, tgt = O.calculate_advantage(r, v, [[False] * 5], 0.97, 0.9, False)
        ra, rt = manual_gae(r[0], v[0], [False] * 5, 0.97, 0.9)
        assert np.allclose(tgt[0], rt, atol=1e-4) and np.allclose(adv[0], ra, atol=1e-4)

    def test_batch_independent(self):
        r = [[1.0, 2.0, 3.0], [10.0, 20.0, 30.0]]
        v = [[0.5] * 4, [1.0] * 4]
        _, tgt = O.calculate_advantage(r, v, np.zeros((2, 3), bool), 0.99, 0.95, False)
        for b in range(2):
            assert np.allclose(tgt[b], manual_gae(r[b], v[b], [False] * 3, 0.99, 0.95)[1], atol=1e-4)

    def test_normalisation(self):
        rng = np.random.default_rng(42)
        adv, _ = O.calculate_advantage(rng.standard_normal((4, 8)), rng.standard_normal((4, 9)),
                                       np.zeros((4, 8), bool), 0.99, 0.95, True)
        assert abs(adv.mean()) < 1e-5 and abs(adv.std() - 1.0) < 0.1

    def test_discount_zero(self):
        r = np.array([[5.0, 10.0, 15.0]], np.float32)
        _, tgt = O.calculate_advantage(r, [[1.0, 2.0, 3.0, 4.0]], np.zeros((1, 3), bool), 0.0, 0.95, False)
        assert np.allclose(tgt, r, atol=1e-5)


class TestValues:  # reference tests/test_values.py
    def test_supports(self):
        s, c = O.calculate_supports(-1.0, 1.0, 10)
        assert s.shape == (1, 11) and c.shape == (10,)
        s, c = O.calculate_supports(-2.0, 3.0, 20)
        assert abs(s[0, 0] + 2.0) < 1e-6 and abs(s[0, -1] - 3.0) < 1e-6
        s, c = O.calculate_supports(0.0, 1.0, 4)
        assert torch.allclose(c, (s[0, :-1] + s[0, 1:]) / 2)

    def test_value_uniform_and_peaked(self):
        _, c = O.calculate_supports(-5.0, 5.0, 51)
        v = O.hlgauss_get_value(torch.zeros(2, 4, 51), c)
        assert v.shape == (2, 4) and torch.allclose(v, torch.zeros(2, 4), atol=0.1)
        lg = torch.full((1, 1, 51), -100.0)
        lg[0, 0, 40] = 100.0
        assert O.hlgauss_get_value(lg, c)[0, 0] > 0

    def test_loss_properties(self):
        args = (-5.0, 5.0, 0.75)
        assert O.hlgauss_get_loss(torch.zeros(2, 4, 51), torch.zeros(2, 4), *args).shape == ()
        tg = torch.zeros(1, 2)
        good = torch.full((1, 2, 51), -10.0); good[:, :, 25] = 10.0
        bad = torch.full((1, 2, 51), -10.0); bad[:, :, 0] = 10.0
        assert O.hlgauss_get_loss(good, tg, *args) < O.hlgauss_get_loss(bad, tg, *args)
        assert torch.isfinite(O.hlgauss_get_loss(torch.zeros(1, 2, 51), torch.tensor([[100.0, -100.0]]), *args))
        lg = torch.zeros(1, 2, 51, requires_grad=True)
        O.hlgauss_get_loss(lg, torch.tensor([[1.0, -1.0]]), *args).backward()
        assert lg.grad.shape == (1, 2, 51) and torch.isfinite(lg.grad).all()

    def test_target_probs_sum_to_one(self):
        p = O.hlgauss_target_probs(torch.tensor([0.3, -9.99, 10.0, 4.2]), -10.0, 10.0, 51, 0.1038)
        assert torch.allclose(p.sum(-1), torch.ones(4), atol=1e-5)
