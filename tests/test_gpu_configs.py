"""Model-level parity on the shapes of every config BASELINE.json names (SURVEY section 8 head + appendix B), not only
the 2-layer one: `return_8_layers` (L=8, T=S=512), `multitask` (H=256, task embeddings fed through `task_ids`, the
unfused rollout chain), `blog_32_layers` (grid_flattened 5x5 encoder, a deep stack).  Layer counts of the 16- and
32-layer configs are cut (4 and 8) so the CPU oracle finishes in seconds; every kernel variant those configs select
(H=256 GEMM shapes, the per-op rollout chain, the flattened encoder) runs.

Reference: TransformerActorCritic.__call__ (network.py:300-343) in both modes, ppo_loss (train.py:162-222).
"""

from dataclasses import replace

import pytest
import torch

pytestmark = pytest.mark.gpu

from jaxrl_b200.config import NAMED, PPOConfig  # noqa: E402

DEV = "cuda"


def _engine(cfg, seed):
    from jaxrl_b200.engine import PolicyEngine
    return PolicyEngine(cfg, DEV, seed=seed, bias_noise=0.05)


# ---------------------------------------------------------------------------------------------- return_8_layers
@pytest.fixture(scope="module")
def eng_return8():
    return _engine(replace(NAMED["return_8_layers"].model, action_dim=6), 21)


def test_return8_decode(eng_return8):
    """L=8, S=512 ring, 72 decode steps through the fused kernels."""
    from _parity import decode_parity
    decode_parity(eng_return8.cfg, eng_return8, steps=72, B=16, seed=31, name="return_8 decode")


def test_return8_train_full_window(eng_return8):
    """L=8 at the full T=512 window (tcgen05 attention fwd/bwd at their bench shape, 4 query tiles per row)."""
    from _parity import train_parity
    train_parity(eng_return8.cfg, eng_return8, NAMED["return_8_layers"].ppo, T=512, B=2, seed=32, name="return_8 T512",
                 logits_allow=4e-2, logits_rel=1.5e-2, grad_rel=4e-2, grad_cos=0.999, value_atol=4e-2)


# ---------------------------------------------------------------------------------------------- multitask (H=256)
@pytest.fixture(scope="module")
def eng_multitask():
    cfg = replace(NAMED["multitask"].model, num_layers=4, max_seq_length=128, action_dim=6)
    eng = _engine(cfg, 22)
    assert cfg.hidden_features == 256 and cfg.task_count == 4
    return eng


def test_multitask_decode(eng_multitask):
    """H=256 + task embeddings: the rollout chain this shape runs, 64 steps."""
    from _parity import decode_parity
    decode_parity(eng_multitask.cfg, eng_multitask, steps=64, B=16, seed=33, name="multitask decode")


def test_multitask_train(eng_multitask):
    from _parity import train_parity
    train_parity(eng_multitask.cfg, eng_multitask, NAMED["multitask"].ppo, T=128, B=4, seed=34, name="multitask T128",
                 logits_allow=5e-3, grad_rel=4e-2, grad_cos=0.999, value_atol=4e-2)


def test_multitask_task_embedding_matters(eng_multitask):
    """task_ids must reach the embedding sum (network.py:308-309): different ids => different logits; the task table
    receives a gradient."""
    from _parity import make_batch
    eng, cfg = eng_multitask, eng_multitask.cfg
    B, T = 2, 128
    batch = make_batch(cfg, T, B, 35)
    ws = eng.alloc_train(B, T)
    d = {k: v.to(DEV) for k, v in batch.items()}
    outs = []
    for shift in (0, 1):
        tid = ((d["task_ids"] + shift) % cfg.task_count).reshape(-1).contiguous()
        _, logits = eng.forward_train(ws, d["obs"], d["last_rewards"].reshape(-1), d["last_actions"].reshape(-1),
                                      action_mask=d["action_mask"], task_ids=tid)
        outs.append(logits.clone())
    assert not torch.equal(outs[0], outs[1])


# ---------------------------------------------------------------------------------------------- blog_32_layers
@pytest.fixture(scope="module")
def eng_blog():
    cfg = replace(NAMED["blog_32_layers"].model, num_layers=8, max_seq_length=128, action_dim=6)
    return _engine(cfg, 23)


def test_blog32_decode(eng_blog):
    from _parity import decode_parity
    decode_parity(eng_blog.cfg, eng_blog, steps=64, B=16, seed=36, name="blog_32 decode")


def test_blog32_train(eng_blog):
    from _parity import train_parity
    train_parity(eng_blog.cfg, eng_blog, NAMED["blog_32_layers"].ppo, T=128, B=4, seed=37, name="blog_32 T128",
                 logits_allow=5e-3, grad_rel=4e-2, grad_cos=0.999, value_atol=4e-2)
