"""CPU: the oracle against the committed fixtures in tests/golden/ (see make_golden.py for what each file holds).

reference_known_answers.npz carries the reference's own test vectors (tests/test_advantage.py, tests/test_values.py);
oracle_*.npz freeze the oracle's outputs so an edit to oracle/oracle.py cannot silently move the parity target."""

import os

import numpy as np
import torch

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def test_gae_reference_known_answers():
    z = load("reference_known_answers.npz")
    for name in ("single", "termination", "loop5", "batch_a", "batch_b", "discount0"):
        disc, lam = z[f"gae_{name}_hyper"]
        adv, tgt = O.calculate_advantage(z[f"gae_{name}_rewards"][None], z[f"gae_{name}_values"][None],
                                         z[f"gae_{name}_terminated"][None], float(disc), float(lam), False)
        np.testing.assert_allclose(tgt[0], z[f"gae_{name}_tgt"], rtol=1e-5, atol=1e-5)   # reference tolerance
        np.testing.assert_allclose(adv[0], z[f"gae_{name}_adv"], rtol=1e-5, atol=1e-5)


def test_supports_reference_vectors():
    z = load("reference_known_answers.npz")
    for name in ("a", "b", "c", "cfg"):
        lo, hi, n = z[f"supports_{name}_args"]
        s, c = O.calculate_supports(float(lo), float(hi), int(n))
        np.testing.assert_allclose(s.reshape(-1).numpy(), z[f"supports_{name}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(c.numpy(), z[f"centers_{name}"], rtol=0, atol=1e-6)


def test_oracle_gae_frozen():
    z = load("oracle_gae.npz")
    disc, lam = z["hyper"]
    adv, tgt = O.calculate_advantage(z["rewards"], z["values"], z["next_terminated"], float(disc), float(lam), False)
    np.testing.assert_array_equal(adv, z["adv"])
    np.testing.assert_array_equal(tgt, z["tgt"])
    adv_n, _ = O.calculate_advantage(z["rewards"], z["values"], z["next_terminated"], float(disc), float(lam), True)
    np.testing.assert_allclose(adv_n, z["adv_norm"], rtol=1e-6, atol=1e-6)


def test_oracle_hlgauss_frozen():
    z = load("oracle_hlgauss.npz")
    vmin, vmax, sigma = (float(v) for v in z["args"])
    logits, targets = torch.from_numpy(z["logits"]), torch.from_numpy(z["targets"])
    _, centers = O.calculate_supports(vmin, vmax, logits.shape[-1])
    np.testing.assert_allclose(O.hlgauss_get_value(logits, centers).numpy(), z["value"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(O.hlgauss_target_probs(targets, vmin, vmax, logits.shape[-1], sigma).numpy(), z["probs"],
                               rtol=1e-5, atol=1e-7)
    lg = logits.clone().requires_grad_(True)
    loss = O.hlgauss_get_loss(lg, targets, vmin, vmax, sigma)
    loss.backward()
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-6)
    np.testing.assert_allclose(lg.grad.numpy(), z["dlogits"], rtol=1e-5, atol=1e-8)


def test_oracle_rope_attention_frozen():
    z = load("oracle_rope_attention.npz")
    bf = lambda k: torch.from_numpy(z[k]).to(torch.bfloat16)
    rope = O.apply_rope(bf("rope_x"), torch.from_numpy(z["rope_pos"]), 32)
    assert torch.equal(rope.float(), torch.from_numpy(z["rope_out"]))
    for kv_len in (1, 17, 48):
        o = O.dot_product_attention(bf("q"), bf("kcache"), bf("vcache"), kv_len=kv_len)
        # a one-ulp bf16 flip is possible across BLAS builds; the fixture was made single-threaded
        np.testing.assert_allclose(o.float().numpy(), z[f"decode_out_len{kv_len}"], rtol=8e-3, atol=1e-3)
    ot = O.dot_product_attention(bf("train_q"), bf("train_k"), bf("train_v"), is_causal=True)
    np.testing.assert_allclose(ot.float().numpy(), z["train_out"], rtol=8e-3, atol=1e-3)


def test_oracle_model_frozen():
    from jaxrl_b200.config import ModelConfig
    from jaxrl_b200.params import ParamStore
    z = load("oracle_model.npz")
    cfg = ModelConfig(num_layers=2, max_seq_length=64, action_dim=6)
    p = ParamStore(cfg, "cpu", seed=int(z["seed"][0]), bias_noise=float(z["bias_noise"][0])).oracle_dict()
    B, T = z["reward"].shape
    ts = O.TimeStep(obs=torch.from_numpy(z["obs"]), time=torch.arange(T, dtype=torch.int32)[None],
                    terminated=torch.zeros(B, T, dtype=torch.bool), last_action=torch.from_numpy(z["last_action"]),
                    reward=torch.from_numpy(z["reward"]), action_mask=torch.from_numpy(z["action_mask"]), task_ids=None)
    vlogits, logits, _ = O.model_forward(p, cfg, ts)
    live = z["action_mask"]
    assert (logits.numpy()[~live] == np.finfo(np.float32).min).all()
    np.testing.assert_allclose(logits.numpy()[live], z["logits"][live], rtol=1e-2, atol=2e-3)
    np.testing.assert_allclose(vlogits.numpy(), z["vlogits"], rtol=1e-2, atol=2e-2)
