"""Regenerates the committed fixtures in this directory:  python tests/golden/make_golden.py

Two kinds of vectors:

* ``reference_known_answers.npz`` - the inputs and expected values the REFERENCE's own tests assert
  (tests/test_advantage.py:38-140 known answers and its ``manual_gae`` loop at :21-35, evaluated here in Python
  double precision exactly as that file does; tests/test_values.py support/centre vectors).  These pin the oracle.
* ``oracle_*.npz`` - outputs of the CPU oracle (oracle/oracle.py) on seeded inputs.  The reference itself cannot be
  imported in the build container (no jax/flax/mapox), so these are NOT reference outputs: they freeze the oracle so
  that a later edit to it cannot silently move the target the CUDA kernels are compared with, and they give the
  ``-m gpu`` tests a fixture that does not depend on the oracle code at all.

Nothing here touches /root/reference at run time.
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402
from jaxrl_b200.config import ModelConfig  # noqa: E402


def manual_gae(rewards, values, terminated, discount, lam):
    """The reference test's loop oracle (tests/test_advantage.py:21-35), python floats."""
    T = len(rewards)
    adv, tgt = [0.0] * T, [0.0] * T
    acc = values[T]
    for t in reversed(range(T)):
        d = 0.0 if terminated[t] else discount
        acc = rewards[t] + d * ((1 - lam) * values[t + 1] + lam * acc)
        tgt[t] = acc
        adv[t] = tgt[t] - values[t]
    return adv, tgt


def reference_known_answers():
    out = {}
    cases = {
        "single": ([1.0], [0.5, 0.8], [False], 0.99, 0.95),
        "termination": ([1.0, 2.0, 3.0], [0.5] * 4, [False, True, False], 0.99, 0.95),
        "loop5": ([1.0, -0.5, 2.0, 0.0, 1.5], [0.3, 0.7, 0.1, 0.9, 0.4, 0.6], [False] * 5, 0.97, 0.9),
        "batch_a": ([1.0, 2.0, 3.0], [0.5] * 4, [False] * 3, 0.99, 0.95),
        "batch_b": ([10.0, 20.0, 30.0], [1.0] * 4, [False] * 3, 0.99, 0.95),
        "discount0": ([5.0, 10.0, 15.0], [1.0, 2.0, 3.0, 4.0], [False] * 3, 0.0, 0.95),
    }
    for name, (r, v, te, disc, lam) in cases.items():
        adv, tgt = manual_gae(r, v, te, disc, lam)
        out[f"gae_{name}_rewards"] = np.asarray(r, np.float32)
        out[f"gae_{name}_values"] = np.asarray(v, np.float32)
        out[f"gae_{name}_terminated"] = np.asarray(te, bool)
        out[f"gae_{name}_hyper"] = np.asarray([disc, lam], np.float64)
        out[f"gae_{name}_adv"] = np.asarray(adv, np.float64)
        out[f"gae_{name}_tgt"] = np.asarray(tgt, np.float64)
    # tests/test_values.py: supports are linspace(min, max, n+1), centres the midpoints
    for name, (lo, hi, n) in {"a": (-1.0, 1.0, 10), "b": (-2.0, 3.0, 20), "c": (0.0, 1.0, 4), "cfg": (-10.0, 10.0, 51)}.items():
        s = np.linspace(lo, hi, n + 1, dtype=np.float64)
        out[f"supports_{name}"] = s
        out[f"centers_{name}"] = (s[:-1] + s[1:]) / 2
        out[f"supports_{name}_args"] = np.asarray([lo, hi, n], np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_known_answers.npz"), **out)


def oracle_gae():
    rng = np.random.default_rng(42)
    B, T = 24, 96
    rewards = (rng.random((B, T)) < 0.05).astype(np.float32) * 0.5 + rng.standard_normal((B, T)).astype(np.float32) * 0.01
    values = rng.standard_normal((B, T + 1)).astype(np.float32)
    term = rng.random((B, T)) < (1.0 / 32)
    term[:, -1] = True
    disc, lam = 0.96609, 0.91237
    adv, tgt = O.calculate_advantage(rewards, values, term, disc, lam, False)
    adv_n, _ = O.calculate_advantage(rewards, values, term, disc, lam, True)
    np.savez_compressed(os.path.join(HERE, "oracle_gae.npz"), rewards=rewards, values=values, next_terminated=term,
                        hyper=np.asarray([disc, lam]), adv=adv, tgt=tgt, adv_norm=adv_n)


def oracle_hlgauss():
    g = torch.Generator().manual_seed(7)
    n, NL, vmin, vmax, sigma = 192, 51, -10.0, 10.0, 0.1038
    logits = torch.randn(n, NL, generator=g) * 2
    targets = torch.randn(n, generator=g) * 4
    targets[:4] = torch.tensor([-25.0, 25.0, -10.0, 10.0])
    supports, centers = O.calculate_supports(vmin, vmax, NL)
    value = O.hlgauss_get_value(logits, centers)
    probs = O.hlgauss_target_probs(targets, vmin, vmax, NL, sigma)
    lg = logits.clone().requires_grad_(True)
    loss = O.hlgauss_get_loss(lg, targets, vmin, vmax, sigma)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "oracle_hlgauss.npz"), logits=logits.numpy(), targets=targets.numpy(),
                        args=np.asarray([vmin, vmax, sigma]), value=value.numpy(), probs=probs.numpy(),
                        loss=loss.detach().numpy(), dlogits=lg.grad.numpy())


def oracle_rope_attention():
    g = torch.Generator().manual_seed(11)
    B, S, Nq, Nkv, D = 6, 48, 4, 1, 32
    # RoPE (positional_embeddings.py:7-27): bf16 in, fp32 math, bf16 out; positions include a wrapped episode
    x = torch.randn(B, 5, Nq, D, generator=g).to(torch.bfloat16)
    pos = torch.tensor([[0, 1, 2, 3, 4], [30, 31, 32, 33, 34], [511, 512, 513, 514, 515], [7, 7, 7, 7, 7],
                        [100, 200, 300, 400, 500], [5, 4, 3, 2, 1]], dtype=torch.int32)
    rope = O.apply_rope(x, pos, D)
    # decode attention (attention.py:138-150): one query vs the first kv_len slots of the ring cache
    q = torch.randn(B, 1, Nq, D, generator=g).to(torch.bfloat16)
    kc = torch.randn(B, S, Nkv, D, generator=g).to(torch.bfloat16)
    vc = torch.randn(B, S, Nkv, D, generator=g).to(torch.bfloat16)
    outs = {}
    for kv_len in (1, 17, 48):
        o = O.dot_product_attention(q, kc, vc, kv_len=kv_len)
        outs[f"decode_out_len{kv_len}"] = o.float().numpy()
    # causal training attention (attention.py:151-169)
    T = 64
    qt = torch.randn(2, T, Nq, D, generator=g).to(torch.bfloat16)
    kt = torch.randn(2, T, Nkv, D, generator=g).to(torch.bfloat16)
    vt = torch.randn(2, T, Nkv, D, generator=g).to(torch.bfloat16)
    ot = O.dot_product_attention(qt, kt, vt, is_causal=True)
    np.savez_compressed(os.path.join(HERE, "oracle_rope_attention.npz"), rope_x=x.float().numpy(), rope_pos=pos.numpy(),
                        rope_out=rope.float().numpy(), q=q.float().numpy(), kcache=kc.float().numpy(),
                        vcache=vc.float().numpy(), train_q=qt.float().numpy(), train_k=kt.float().numpy(),
                        train_v=vt.float().numpy(), train_out=ot.float().numpy(), **outs)


def oracle_model():
    """Tiny trunk forward (network.py:300-343): logits, value logits and value on a seeded 2-layer model."""
    from jaxrl_b200.params import ParamStore
    cfg = ModelConfig(num_layers=2, max_seq_length=64, action_dim=6)
    store = ParamStore(cfg, "cpu", seed=3, bias_noise=0.05)
    p = store.oracle_dict()
    g = torch.Generator().manual_seed(19)
    B, T = 2, 64      # the training attention kernels take windows that are multiples of 64
    vw, vh, K = cfg.obs_shape
    obs = torch.stack([torch.randint(0, n, (B, T, vw, vh), generator=g) for n in cfg.obs_max_value], dim=-1).to(torch.uint8)
    reward = torch.randn(B, T, generator=g) * 0.1
    last_action = torch.randint(0, cfg.action_dim, (B, T), generator=g).to(torch.int32)
    last_action[:, 0] = -1
    mask = torch.rand(B, T, cfg.action_dim, generator=g) < 0.8
    mask[..., 0] = True
    time = torch.arange(T, dtype=torch.int32)[None]
    ts = O.TimeStep(obs=obs, time=time, terminated=torch.zeros(B, T, dtype=torch.bool), last_action=last_action,
                    reward=reward, action_mask=mask, task_ids=None)
    vlogits, logits, _ = O.model_forward(p, cfg, ts)
    _, centers = O.calculate_supports(cfg.value_min, cfg.value_max, cfg.value_n_logits)
    value = O.hlgauss_get_value(vlogits, centers)
    np.savez_compressed(os.path.join(HERE, "oracle_model.npz"), obs=obs.numpy(), reward=reward.numpy(),
                        last_action=last_action.numpy(), action_mask=mask.numpy(), logits=logits.float().numpy(),
                        vlogits=vlogits.float().numpy(), value=value.float().numpy(), seed=np.asarray([3]),
                        bias_noise=np.asarray([0.05]))


if __name__ == "__main__":
    torch.set_num_threads(1)
    reference_known_answers()
    oracle_gae()
    oracle_hlgauss()
    oracle_rope_attention()
    oracle_model()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
