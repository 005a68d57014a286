"""The reference's own tests (tests/test_attention.py, test_rope.py, test_advantage.py, test_values.py) re-expressed
against the B200 mirror modules (same class / method names, same arguments).  The reference tests run fp32 modules
with atol 1e-5; the B200 path is bf16 compute, so activations are compared at bf16 tolerance while the fp32 kernels
(GAE, HL-Gauss) keep the reference's own tolerances."""

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
BF16 = torch.bfloat16


def make_attention(d_model=128, head_dim=32, num_heads=4, num_kv_heads=4, max_seq_length=64, **kw):
    from jaxrl_b200.model.attention import AttentionBlock
    gen = torch.Generator().manual_seed(0)
    return AttentionBlock(d_model=d_model, head_dim=head_dim, num_heads=num_heads, num_kv_heads=num_kv_heads,
                          max_seq_length=max_seq_length, rngs=0,
                          kernel_init=lambda shape: torch.randn(shape, generator=gen) * 0.1, **kw)


class TestAttentionBlock:      # reference tests/test_attention.py:35-80
    def test_output_shape(self):
        attn = make_attention()
        out, _ = attn(torch.ones(2, 64, 128, dtype=BF16, device=DEV), torch.arange(64, device=DEV)[None].repeat(2, 1))
        assert out.shape == (2, 64, 128)

    def test_causal_masking(self):
        attn = make_attention()
        x = torch.randn(1, 64, 128, generator=torch.Generator().manual_seed(0)).to(BF16).to(DEV)
        pos = torch.arange(64, device=DEV)[None]
        out1, _ = attn(x, pos)
        x2 = x.clone()
        x2[0, 63, :] = 99.0
        out2, _ = attn(x2, pos)
        assert torch.equal(out1[0, :63], out2[0, :63])           # earlier positions cannot see position 63
        assert not torch.allclose(out1[0, 63].float(), out2[0, 63].float(), atol=1e-3)

    @pytest.mark.parametrize("kv", [2, 1])
    def test_gqa_mqa(self, kv):
        attn = make_attention(num_heads=4, num_kv_heads=kv)
        out, _ = attn(torch.ones(2, 64, 128, dtype=BF16, device=DEV), torch.arange(64, device=DEV)[None].repeat(2, 1))
        assert out.shape == (2, 64, 128) and torch.isfinite(out.float()).all()

    def test_indivisible_d_model_raises(self):
        from jaxrl_b200.model.attention import AttentionBlock
        with pytest.raises(ValueError, match="must be divisible"):
            AttentionBlock(d_model=130, head_dim=32, num_heads=4, num_kv_heads=1, max_seq_length=8)


class TestKVCache:             # reference tests/test_attention.py:83-136
    def test_initialize_carry_shape(self):
        attn = make_attention(max_seq_length=16, num_kv_heads=2)
        cache = attn.initialize_carry(batch_size=3)
        assert cache.key.shape == (3, 16, 2, 32) and cache.value.shape == (3, 16, 2, 32)
        assert (cache.key == 0).all()

    def test_stores_at_correct_position(self):
        attn = make_attention(max_seq_length=8)
        cache = attn.initialize_carry(batch_size=1)
        upd = attn.update_kv_cache(cache, torch.tensor([[3]]), torch.ones(1, 1, 4, 32, device=DEV),
                                   torch.ones(1, 1, 4, 32, device=DEV) * 2.0)
        assert (upd.key[0, 3] == 1.0).all() and (upd.value[0, 3] == 2.0).all() and (upd.key[0, 0] == 0).all()

    def test_wraps_around(self):
        attn = make_attention(max_seq_length=4)
        cache = attn.initialize_carry(batch_size=1)
        upd = attn.update_kv_cache(cache, torch.tensor([[6]]), torch.ones(1, 1, 4, 32, device=DEV) * 5.0,
                                   torch.ones(1, 1, 4, 32, device=DEV) * 5.0)
        assert (upd.key[0, 2] == 5.0).all() and (upd.key[0, 0] == 0).all()

    def test_call_writes_ring_slot(self):
        attn = make_attention(max_seq_length=4, num_kv_heads=1)
        cache = attn.initialize_carry(batch_size=2)
        x = torch.randn(2, 1, 128, generator=torch.Generator().manual_seed(2)).to(BF16).to(DEV)
        _, cache = attn(x, torch.full((2, 1), 6, device=DEV), cache)
        assert cache.key[:, 2].abs().sum() > 0 and (cache.key[:, [0, 1, 3]] == 0).all()

    def test_inference_matches_training(self):
        """Every decode step equals the matching position of the full causal pass (stronger than the reference's
        first-step check, tests/test_attention.py:123-136)."""
        attn = make_attention(max_seq_length=64, num_kv_heads=1)
        x = torch.randn(2, 64, 128, generator=torch.Generator().manual_seed(1)).to(BF16).to(DEV)
        pos = torch.arange(64, device=DEV)[None].repeat(2, 1)
        out_train, _ = attn(x, pos, kv_cache=None)
        cache = attn.initialize_carry(batch_size=2)
        for t in range(64):
            o, cache = attn(x[:, t:t + 1].contiguous(), pos[:, t:t + 1].contiguous(), cache)
            torch.testing.assert_close(o[:, 0].float(), out_train[:, t].float(), rtol=2e-2, atol=2e-2)


class TestApplyRope:           # reference tests/test_rope.py
    def _rope(self, x, pos, d):
        from jaxrl_b200.model.positional_embeddings import apply_rope
        return apply_rope(x.to(BF16).to(DEV), pos.to(DEV), head_dim=d)

    def test_shape_and_dtype(self):
        out = self._rope(torch.randn(2, 8, 4, 32), torch.arange(8)[None].repeat(2, 1), 32)
        assert out.shape == (2, 8, 4, 32) and out.dtype == BF16

    def test_preserves_norm(self):
        x = torch.randn(2, 8, 4, 32, generator=torch.Generator().manual_seed(1)).to(BF16)
        out = self._rope(x, torch.arange(8)[None].repeat(2, 1), 32)
        torch.testing.assert_close(out.float().norm(dim=-1).cpu(), x.float().norm(dim=-1), rtol=1e-2, atol=1e-2)

    def test_position_zero_is_identity(self):
        x = torch.randn(1, 1, 2, 32, generator=torch.Generator().manual_seed(2)).to(BF16)
        assert torch.equal(self._rope(x, torch.zeros(1, 1, dtype=torch.int32), 32).cpu(), x)

    def test_positions_differ(self):
        out = self._rope(torch.ones(1, 2, 1, 32), torch.tensor([[0, 5]]), 32)
        assert not torch.allclose(out[0, 0].float(), out[0, 1].float(), atol=1e-3)

    def test_relative_position(self):
        q = torch.randn(1, 1, 1, 32, generator=torch.Generator().manual_seed(3))
        k = torch.randn(1, 1, 1, 32, generator=torch.Generator().manual_seed(4))
        d1 = (self._rope(q, torch.tensor([[2]]), 32).float() * self._rope(k, torch.tensor([[5]]), 32).float()).sum()
        d2 = (self._rope(q, torch.tensor([[10]]), 32).float() * self._rope(k, torch.tensor([[13]]), 32).float()).sum()
        assert abs(d1.item() - d2.item()) < 0.1      # bf16 inputs (reference: fp32, 1e-4)


def make_rollout(batch_size=2, trajectory_length=4):
    from jaxrl_b200.rollout import ActionSpec, ObservationSpec, Rollout
    return Rollout(batch_size, trajectory_length, ObservationSpec((3, 3, 2), torch.uint8, (5, 5)), ActionSpec(4))


def manual_gae(rewards, values, terminated, discount, lam):     # reference tests/test_advantage.py:21-35
    T = len(rewards)
    adv, tgt = [0.0] * T, [0.0] * T
    acc = values[T]
    for t in reversed(range(T)):
        d = 0.0 if terminated[t] else discount
        acc = rewards[t] + d * ((1 - lam) * values[t + 1] + lam * acc)
        tgt[t] = acc
        adv[t] = tgt[t] - values[t]
    return adv, tgt


class TestCalculateAdvantage:  # reference tests/test_advantage.py
    def _run(self, rewards, values, term, discount, lam, norm=False):
        B, T = len(rewards), len(rewards[0])
        ro = make_rollout(B, T)
        st = ro.create_state()._replace(rewards=torch.tensor(rewards, device=DEV), values=torch.tensor(values, device=DEV),
                                        next_terminated=torch.tensor(term, device=DEV))
        return ro.calculate_advantage(st, discount=discount, gae_lambda=lam, norm_adv=norm)

    def test_single_step(self):
        r = self._run([[1.0]], [[0.5, 0.8]], [[False]], 0.99, 0.95)
        assert abs(r.targets[0, 0].item() - (1.0 + 0.99 * 0.8)) < 1e-5
        assert abs(r.advantages[0, 0].item() - (1.0 + 0.99 * 0.8 - 0.5)) < 1e-5

    def test_termination_zeros_discount(self):
        rw, v, te = [[1.0, 2.0, 3.0]], [[0.5, 0.5, 0.5, 0.5]], [[False, True, False]]
        r = self._run(rw, v, te, 0.99, 0.95)
        _, ref = manual_gae(rw[0], v[0], te[0], 0.99, 0.95)
        for t in range(3):
            assert abs(r.targets[0, t].item() - ref[t]) < 1e-5

    def test_matches_loop(self):
        rw, v = [[1.0, -0.5, 2.0, 0.0, 1.5]], [[0.3, 0.7, 0.1, 0.9, 0.4, 0.6]]
        r = self._run(rw, v, [[False] * 5], 0.97, 0.9)
        ra, rt = manual_gae(rw[0], v[0], [False] * 5, 0.97, 0.9)
        for t in range(5):
            assert abs(r.targets[0, t].item() - rt[t]) < 1e-4 and abs(r.advantages[0, t].item() - ra[t]) < 1e-4

    def test_batch_independent(self):
        rw, v = [[1.0, 2.0, 3.0], [10.0, 20.0, 30.0]], [[0.5] * 4, [1.0] * 4]
        r = self._run(rw, v, [[False] * 3] * 2, 0.99, 0.95)
        for b in range(2):
            _, rt = manual_gae(rw[b], v[b], [False] * 3, 0.99, 0.95)
            for t in range(3):
                assert abs(r.targets[b, t].item() - rt[t]) < 1e-4

    def test_normalisation(self):
        g = torch.Generator().manual_seed(42)
        r = self._run(torch.randn(4, 8, generator=g).tolist(), torch.randn(4, 9, generator=g).tolist(),
                      [[False] * 8] * 4, 0.99, 0.95, norm=True)
        assert abs(r.advantages.mean().item()) < 1e-5 and abs(r.advantages.std(unbiased=False).item() - 1.0) < 0.1

    def test_discount_zero(self):
        rw = [[5.0, 10.0, 15.0]]
        r = self._run(rw, [[1.0, 2.0, 3.0, 4.0]], [[False] * 3], 0.0, 0.95)
        assert torch.allclose(r.targets.cpu(), torch.tensor(rw), atol=1e-5)


class TestHlGaussValue:        # reference tests/test_values.py
    @pytest.fixture
    def value_head(self):
        from jaxrl_b200.values import HlGaussConfig, HlGaussValue
        return HlGaussValue(32, HlGaussConfig(min=-5.0, max=5.0, n_logits=51, sigma=0.75), rngs=0)

    def test_supports(self):
        from jaxrl_b200.values import HlGaussConfig, calculate_supports
        s, c = calculate_supports(HlGaussConfig(min=-1.0, max=1.0, n_logits=10, sigma=0.5))
        assert s.shape == (1, 11) and c.shape == (10,)
        s, c = calculate_supports(HlGaussConfig(min=-2.0, max=3.0, n_logits=20, sigma=0.5))
        assert abs(s[0, 0] + 2.0) < 1e-6 and abs(s[0, -1] - 3.0) < 1e-6
        s, c = calculate_supports(HlGaussConfig(min=0.0, max=1.0, n_logits=4, sigma=0.5))
        assert torch.allclose(c, (s[0, :-1] + s[0, 1:]) / 2)

    def test_get_value(self, value_head):
        v = value_head.get_value(torch.zeros(2, 4, 51, device=DEV))
        assert v.shape == (2, 4) and torch.allclose(v, torch.zeros_like(v), atol=0.1)
        lg = torch.full((1, 1, 51), -100.0, device=DEV)
        lg[0, 0, 40] = 100.0
        assert value_head.get_value(lg)[0, 0] > 0

    def test_call_shape(self, value_head):
        out = value_head(torch.randn(3, 5, 32, device=DEV).to(BF16))
        assert out.shape == (3, 5, 51) and out.dtype == torch.float32

    def test_loss(self, value_head):
        assert value_head.get_loss(torch.zeros(2, 4, 51, device=DEV), torch.zeros(2, 4, device=DEV)).shape == ()
        tg = torch.zeros(1, 2, device=DEV)
        good = torch.full((1, 2, 51), -10.0, device=DEV); good[:, :, 25] = 10.0
        bad = torch.full((1, 2, 51), -10.0, device=DEV); bad[:, :, 0] = 10.0
        assert value_head.get_loss(good, tg) < value_head.get_loss(bad, tg)
        assert torch.isfinite(value_head.get_loss(torch.zeros(1, 2, 51, device=DEV), torch.tensor([[100.0, -100.0]], device=DEV)))
        _, grad = value_head.get_loss_and_grad(torch.zeros(1, 2, 51, device=DEV), torch.tensor([[1.0, -1.0]], device=DEV))
        assert grad.shape == (1, 2, 51) and torch.isfinite(grad).all()


def test_glu_block_matches_oracle():
    from jaxrl_b200.model.feed_forward import GLUBlock
    from oracle import oracle as O
    blk = GLUBlock(128, 768, rngs=3)
    x = torch.randn(4, 9, 128, generator=torch.Generator().manual_seed(5)).to(BF16)
    ref = O.glu_block({"f.up_proj.kernel": blk.up_proj.cpu(), "f.up_gate.kernel": blk.up_gate.cpu(),
                       "f.down_proj.kernel": blk.down_proj.cpu()}, "f.", x)
    out = blk(x.to(DEV))
    from _util import assert_close_bf16
    assert_close_bf16(out, ref, "GLUBlock mirror", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=1e-3, max_err_frac=0.1)
