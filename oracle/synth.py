"""Self-contained workload description for the CPU arm: config literals, parameter shapes / initialisers and the
synthetic PPO update (rollout -> GAE -> minibatch fwd/bwd) built from the restatement in `oracle.py`.

TEST / BASELINE INFRASTRUCTURE ONLY (see the package docstring): `bench.py --impl reference` and `bench.py`'s
`cpu_baseline` leg run this file; it imports nothing from `jaxrl_b200`, so the CPU arm never loads the CUDA library.
`tests/test_oracle_synth.py` checks that the literals below agree with `jaxrl_b200.config.NAMED` and
`jaxrl_b200.params.param_specs`.

Shapes follow the reference's config files (config/return_2_layers.json etc.);
initialisers restate network.py:39-52,190-193 (glorot-uniform layer kernels, N(0,1)*0.01 embedding tables, flax
defaults elsewhere).
"""

from __future__ import annotations

import math
import os
import statistics
import time
from types import SimpleNamespace

import torch

from . import oracle as O


def _model(**kw):
    base = dict(hidden_features=128, num_layers=2, num_heads=4, num_kv_heads=1, head_dim=32,
                rope_max_wavelength=10_000.0, ffn_size=768, value_hidden_dim=768, value_min=-10.0, value_max=10.0,
                value_n_logits=51, value_sigma=0.10381270166884848, obs_type="grid_cnn",
                cnn_kernels=((3, 3), (3, 3), (3, 3)), cnn_strides=((2, 2), (1, 1), (1, 1)), cnn_channels=(16, 32),
                obs_shape=(11, 11, 2), obs_max_value=(5, 5), action_dim=8, task_count=1, max_seq_length=512)
    base.update(kw)
    return SimpleNamespace(**base)


def _ppo(**kw):
    base = dict(epoch_count=1, minibatch_count=16, vf_coef=0.5392535826611433, entropy_coef=0.001,
                entropy_coef_end=0.0, vf_clip=0.2827893540340816, discount=0.9660902557298033,
                gae_lambda=0.9123716252262187, normalize_advantage=True)
    base.update(kw)
    return SimpleNamespace(**base)


# name -> (model config, ppo hyper-parameters, agents B of the full-size run)
WORKLOADS = {
    "test": (_model(max_seq_length=64), _ppo(minibatch_count=2), 16),
    "return_2_layers": (_model(), _ppo(), 4096),
    "return_8_layers": (_model(num_layers=8), _ppo(), 4096),
    "multitask": (_model(hidden_features=256, num_layers=16, task_count=4),
                  _ppo(minibatch_count=32, discount=0.99, gae_lambda=0.91), 4096),
    "blog_32_layers": (_model(num_layers=32, obs_type="grid_flattened", obs_shape=(5, 5), obs_max_value=(8,)),
                       _ppo(normalize_advantage=False, entropy_coef=8.270466882977551e-05), 4096),
}


def param_specs(cfg):
    """(name, shape, initialiser kind) for every nnx.Param of TransformerActorCritic (network.py:206-295)."""
    H, D, Nq, Nkv, F = cfg.hidden_features, cfg.head_dim, cfg.num_heads, cfg.num_kv_heads, cfg.ffn_size
    specs = [("reward_encoder.kernel", (1, H), "lecun"), ("reward_encoder.bias", (H,), "zeros"),
             ("action_embedder.embedding_table", (cfg.action_dim, H), "embed")]
    C = sum(cfg.obs_max_value)
    if cfg.obs_type == "grid_cnn":
        cin = C
        for j, ((kh, kw), ch) in enumerate(zip(cfg.cnn_kernels, list(cfg.cnn_channels) + [H])):
            specs += [(f"obs_encoder.layers.{j}.kernel", (kh, kw, cin, ch), "lecun"),
                      (f"obs_encoder.layers.{j}.bias", (ch,), "zeros")]
            cin = ch
    elif cfg.obs_type == "grid_flattened":
        specs += [("obs_encoder.embedding.kernel", (C, 4), "lecun"), ("obs_encoder.embedding.bias", (4,), "zeros"),
                  ("obs_encoder.dense.kernel", (4 * cfg.obs_shape[0] * cfg.obs_shape[1], H), "lecun"),
                  ("obs_encoder.dense.bias", (H,), "zeros")]
    else:
        raise ValueError(cfg.obs_type)
    if cfg.task_count > 1:
        specs.append(("task_embedder.embedding_table", (cfg.task_count, H), "embed"))
    for i in range(cfg.num_layers):
        pre = f"layers.{i}."
        specs += [(pre + "history_norm.scale", (H,), "ones"),
                  (pre + "history.key_proj.kernel", (H, Nkv, D), "glorot"),
                  (pre + "history.value_proj.kernel", (H, Nkv, D), "glorot"),
                  (pre + "history.query_proj.kernel", (H, Nq, D), "glorot"),
                  (pre + "history.out.kernel", (Nq, D, H), "glorot_out"),
                  (pre + "ffn_norm.scale", (H,), "ones"),
                  (pre + "ffn.up_proj.kernel", (H, F), "glorot"),
                  (pre + "ffn.up_gate.kernel", (H, F), "glorot"),
                  (pre + "ffn.down_proj.kernel", (F, H), "glorot")]
    specs.append(("output_norm.scale", (H,), "ones"))
    vin = H
    if cfg.value_hidden_dim is not None:
        specs += [("value_mlp.kernel", (H, cfg.value_hidden_dim), "lecun"), ("value_mlp.bias", (cfg.value_hidden_dim,), "zeros")]
        vin = cfg.value_hidden_dim
    specs += [("value_head.dense.kernel", (vin, cfg.value_n_logits), "lecun"),
              ("value_head.dense.bias", (cfg.value_n_logits,), "zeros")]
    return specs


def init_tensor(shape, kind: str, gen: torch.Generator) -> torch.Tensor:
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "ones":
        return torch.ones(shape)
    if kind == "embed":
        return torch.randn(shape, generator=gen) * 0.01
    n = math.prod(shape)
    if kind == "glorot":          # (in, *out)
        fan_in, fan_out = shape[0], n // shape[0]
    else:                         # glorot_out / lecun: (*receptive, in, out)
        fan_in, fan_out = n // shape[-1], shape[-1]
    if kind.startswith("glorot"):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=gen) * 2 - 1) * lim
    return torch.randn(shape, generator=gen) * math.sqrt(1.0 / fan_in)


def init_params(cfg, seed: int = 0) -> dict:
    gen = torch.Generator().manual_seed(seed)
    return {name: init_tensor(shape, kind, gen) for name, shape, kind in param_specs(cfg)}


def synthetic_inputs(cfg, agents: int, seed: int = 0):
    """Observation / reward / Gumbel-noise streams of the named env shapes (SURVEY 8d "synthetic inputs")."""
    gen = torch.Generator().manual_seed(seed)
    T = cfg.max_seq_length
    vw, vh = cfg.obs_shape[0], cfg.obs_shape[1]
    comps = [torch.randint(0, n, (T + 1, agents, vw, vh), generator=gen) for n in cfg.obs_max_value]
    obs = (comps[0] if cfg.obs_type == "grid_flattened" else torch.stack(comps, -1)).to(torch.uint8)
    reward = (torch.rand(T + 1, agents, generator=gen) < 0.02).float() * 0.05
    u = torch.rand(T, agents, cfg.action_dim, generator=gen).clamp(1e-6, 1 - 1e-6)
    gumbel = -torch.log(-torch.log(u))
    task_ids = torch.randint(0, cfg.task_count, (agents,), generator=gen).to(torch.int32) if cfg.task_count > 1 else None
    return obs, reward, gumbel, task_ids


def ppo_update(p: dict, cfg, hyp, obs, reward, gumbel, task_ids=None):
    """One PPO update of the reference's train() (train.py:225-309) without the optimizer: rollout of T decode steps,
    GAE, minibatch_count x (ppo_loss forward + backward).  Returns the gradients of the last minibatch."""
    agents = obs.shape[1]
    T = cfg.max_seq_length
    with torch.no_grad():
        ro = O.rollout_decode(p, cfg, obs, reward, gumbel, task_ids=task_ids)
    term = torch.zeros(agents, T, dtype=torch.bool)
    term[:, -1] = True
    adv, tgt = O.calculate_advantage(ro["rewards"].numpy(), ro["values"].numpy(), term.numpy(), hyp.discount,
                                     hyp.gae_lambda, hyp.normalize_advantage)
    adv, tgt = torch.from_numpy(adv), torch.from_numpy(tgt)
    m = max(1, min(hyp.minibatch_count, agents))
    bm = agents // m
    pg = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    bobs = obs[:T].transpose(0, 1)
    for i in range(m):
        sl = slice(i * bm, (i + 1) * bm)
        batch = dict(obs=bobs[sl], action_mask=None, actions=ro["actions"][sl], last_actions=ro["last_actions"][sl],
                     last_rewards=ro["last_rewards"][sl], terminated=torch.zeros(bm, T, dtype=torch.bool),
                     log_prob=ro["log_prob"][sl], advantages=adv[sl], targets=tgt[sl])
        if task_ids is not None:
            batch["task_ids"] = task_ids[sl][:, None].expand(bm, T)
        total, _ = O.ppo_loss(pg, cfg, hyp, batch, 0.0)
        total.backward()
    return {k: v.grad for k, v in pg.items()}, m


def time_cpu_update(name: str, agents: int, steps: int = 1, warmup: int = 0, threads: int | None = None):
    """Times `steps` PPO updates of workload `name` on a bounded sample of `agents` agents with all host threads.
    -> (cpu_baseline dict for bench.py, seconds per update)."""
    cfg, hyp, _ = WORKLOADS[name]
    torch.set_num_threads(threads or os.cpu_count() or 1)
    p = init_params(cfg, 0)
    obs, reward, gumbel, task_ids = synthetic_inputs(cfg, agents, 0)
    times = []
    m = 1
    for _ in range(warmup + steps):
        t0 = time.perf_counter()
        _, m = ppo_update(p, cfg, hyp, obs, reward, gumbel, task_ids)
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times[warmup:])
    T = cfg.max_seq_length
    return dict(value=agents * T / dt, unit="env-steps/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{agents} agents x {T} steps x {cfg.num_layers} layers, {m} minibatches, torch-CPU oracle "
                       f"(restatement of the reference; JAX CPU backend not installable here), {dt:.2f} s per update"), dt
