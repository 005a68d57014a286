"""GPU: the CUDA path (through the C ABI) against the committed fixtures in tests/golden/ - no oracle code is
imported here, so these comparisons stand even if oracle/ changes.  Tolerances are BASELINE.json's north star:
indices / sequential fp32 bit-exact, fp32 <= 1e-5, bf16 <= 1e-2 relative (with an rms-scaled floor near zero)."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import ULP2, assert_close_bf16, rope_timescale  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.fixture(scope="module")
def ops():
    from jaxrl_b200 import ops as _ops
    return _ops


def test_gae_reference_known_answers(ops):
    z = load("reference_known_answers.npz")
    for name in ("single", "termination", "loop5", "batch_a", "batch_b", "discount0"):
        disc, lam = z[f"gae_{name}_hyper"]
        adv, tgt, _ = ops.gae(torch.from_numpy(z[f"gae_{name}_rewards"])[None].to(DEV),
                              torch.from_numpy(z[f"gae_{name}_values"])[None].to(DEV),
                              torch.from_numpy(z[f"gae_{name}_terminated"])[None].to(DEV), float(disc), float(lam))
        np.testing.assert_allclose(tgt.cpu().numpy()[0], z[f"gae_{name}_tgt"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(adv.cpu().numpy()[0], z[f"gae_{name}_adv"], rtol=1e-5, atol=1e-5)


def test_gae_fixture_bit_exact(ops):
    z = load("oracle_gae.npz")
    disc, lam = z["hyper"]
    adv, tgt, stats = ops.gae(torch.from_numpy(z["rewards"]).to(DEV), torch.from_numpy(z["values"]).to(DEV),
                              torch.from_numpy(z["next_terminated"]).to(DEV), float(disc), float(lam))
    np.testing.assert_array_equal(adv.cpu().numpy(), z["adv"])
    np.testing.assert_array_equal(tgt.cpu().numpy(), z["tgt"])
    ops.adv_normalize(adv, stats)
    np.testing.assert_allclose(adv.cpu().numpy(), z["adv_norm"], rtol=1e-5, atol=1e-5)


def test_hlgauss_fixture(ops):
    z = load("oracle_hlgauss.npz")
    vmin, vmax, sigma = (float(v) for v in z["args"])
    logits, targets = torch.from_numpy(z["logits"]).to(DEV), torch.from_numpy(z["targets"]).to(DEV)
    NL = logits.shape[-1]
    supports = torch.linspace(vmin, vmax, NL + 1, dtype=torch.float32)
    centers = (supports[:-1] + supports[1:]) / 2
    val = ops.hlgauss_value(logits, centers.to(DEV))
    np.testing.assert_allclose(val.cpu().numpy(), z["value"], rtol=1e-5, atol=1e-5)
    loss, dl = ops.hlgauss_loss(logits, targets, supports.to(DEV), vmin, vmax, sigma, 1.0)
    np.testing.assert_allclose(loss.cpu().numpy()[0], z["loss"], rtol=2e-5)
    np.testing.assert_allclose(dl.cpu().numpy(), z["dlogits"], rtol=1e-4, atol=1e-7)


def test_rope_fixture(ops):
    z = load("oracle_rope_attention.npz")
    x = torch.from_numpy(z["rope_x"]).to(BF16)              # (B, 5, Nq, D)
    B, T, Nq, D = x.shape
    xd = x.reshape(B * T, Nq * D).clone().to(DEV)
    pos = torch.from_numpy(z["rope_pos"]).reshape(-1).to(DEV)
    ops.rope_(xd, Nq * D, pos, rope_timescale(D).to(DEV), B * T, Nq, D, inverse=False)
    assert_close_bf16(xd.reshape(B, T, Nq, D), torch.from_numpy(z["rope_out"]), "rope vs fixture")


@pytest.mark.parametrize("impl", [0, 1, 2])
def test_decode_attention_fixture(ops, impl):
    from jaxrl_b200._lib import lib
    z = load("oracle_rope_attention.npz")
    q = torch.from_numpy(z["q"]).to(BF16)                    # (B, 1, Nq, D)
    kc, vc = torch.from_numpy(z["kcache"]).to(BF16).to(DEV), torch.from_numpy(z["vcache"]).to(BF16).to(DEV)
    B, _, Nq, D = q.shape
    lib.mpx_debug_set(1, impl)
    try:
        for kv_len in (1, 17, 48):
            pos = torch.full((B,), kv_len - 1, dtype=torch.int32, device=DEV)
            out = ops.attn_decode(q.reshape(B, Nq * D).to(DEV), kc, vc, pos, Nq, kc.shape[2], D)
            torch.cuda.synchronize()
            assert_close_bf16(out.reshape(B, Nq, D), torch.from_numpy(z[f"decode_out_len{kv_len}"])[:, 0],
                              f"decode attention fixture impl={impl} kv_len={kv_len}", rtol=ULP2 if impl == 0 else 1e-2,
                              abs_floor_frac=1e-2 if impl == 0 else 4e-3)
    finally:
        lib.mpx_debug_set(1, 0)


@pytest.mark.parametrize("impl", [0, 1])
def test_train_attention_fixture(ops, impl):
    from jaxrl_b200._lib import lib
    z = load("oracle_rope_attention.npz")
    q, k, v = (torch.from_numpy(z[n]).to(BF16) for n in ("train_q", "train_k", "train_v"))
    B, T, Nq, D = q.shape
    Nkv = k.shape[2]
    qkv = torch.cat([q.reshape(B * T, -1), k.reshape(B * T, -1), v.reshape(B * T, -1)], dim=1).contiguous().to(DEV)
    QD, KD = Nq * D, Nkv * D
    ld = QD + 2 * KD
    lib.mpx_debug_set(2, impl)
    try:
        o, _ = ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], B, T, Nq, Nkv, D, ld, ld, ld)
        torch.cuda.synchronize()
    finally:
        lib.mpx_debug_set(2, 0)
    assert_close_bf16(o.reshape(B, T, Nq, D), torch.from_numpy(z["train_out"]), f"train attention fixture impl={impl}",
                      rtol=ULP2, abs_floor_frac=1e-2)


def test_model_forward_fixture(ops):
    from jaxrl_b200.config import ModelConfig
    from jaxrl_b200.engine import PolicyEngine
    z = load("oracle_model.npz")
    cfg = ModelConfig(num_layers=2, max_seq_length=64, action_dim=6)
    eng = PolicyEngine(cfg, DEV, seed=int(z["seed"][0]), bias_noise=float(z["bias_noise"][0]))
    B, T = z["reward"].shape
    ws = eng.alloc_train(B, T)
    vlogits, logits = eng.forward_train(ws, torch.from_numpy(z["obs"]).to(DEV),
                                        torch.from_numpy(z["reward"]).reshape(-1).to(DEV),
                                        torch.from_numpy(z["last_action"]).reshape(-1).to(DEV),
                                        action_mask=torch.from_numpy(z["action_mask"]).to(DEV))
    A = cfg.action_dim
    lg = logits.cpu().reshape(B, T, A)
    live = torch.from_numpy(z["action_mask"])
    assert (lg[~live] == torch.finfo(torch.float32).min).all()
    assert_close_bf16(lg[live], torch.from_numpy(z["logits"])[live], "fixture logits", rtol=1e-2, abs_floor_frac=2e-2,
                      allow_frac=5e-3, max_err_frac=6e-2)
    assert_close_bf16(vlogits.cpu().reshape(B, T, -1), torch.from_numpy(z["vlogits"]), "fixture vlogits", rtol=1e-2,
                      abs_floor_frac=2e-2, allow_frac=5e-3, max_err_frac=6e-2)
    val = ops.hlgauss_value(vlogits, eng.centers).cpu().reshape(B, T)
    torch.testing.assert_close(val, torch.from_numpy(z["value"]), rtol=1e-2, atol=2e-2)
