"""Torch-CPU restatement of the mapox_trainer policy hot path (see package docstring).

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Arrays that are bf16 in the reference are torch.bfloat16
tensors here; arithmetic is done in fp32 with a single rounding where the
JAX/flax op has a single rounding (SURVEY.md section 8 "numerics contract").

Parameters are a flat ``dict[str, torch.Tensor]`` (fp32) keyed by the
reference's nnx pytree paths, e.g. ``layers.0.history.query_proj.kernel``.
The model config is duck-typed (any object with the attributes used below;
``jaxrl_b200.config.ModelConfig`` in the tests).
"""

from __future__ import annotations

import math
from typing import NamedTuple, Optional

import numpy as np
import torch

BF16 = torch.bfloat16
F32 = torch.float32


# --------------------------------------------------------------------------- helpers
def _f(x: torch.Tensor) -> torch.Tensor:
    return x.to(F32)


def _b(x: torch.Tensor) -> torch.Tensor:
    return x.to(BF16)


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """jax.nn.gelu(approximate=True) (network.py:30,338): fp32 math, one rounding to x.dtype."""
    xf = _f(x)
    c = math.sqrt(2.0 / math.pi)
    y = 0.5 * xf * (1.0 + torch.tanh(c * (xf + 0.044715 * xf * xf * xf)))
    return y.to(x.dtype)


def linear(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor] = None,
           dtype: Optional[torch.dtype] = BF16) -> torch.Tensor:
    """flax nnx.Linear / LinearGeneral (contract #1).

    dtype=bf16: inputs, kernel, bias cast to bf16; fp32 accumulate; one rounding to bf16;
    bias added in bf16 (second rounding).  dtype=None: promote to fp32 (values.py:34).
    ``kernel`` is (..in_dims.., ..out_dims..); contraction is over the leading ``k`` dims of the
    kernel that match x's trailing dims (the caller flattens).
    """
    if dtype is None:
        y = _f(x) @ _f(kernel)
        if bias is not None:
            y = y + _f(bias)
        return y
    y = (_f(x.to(dtype)) @ _f(kernel.to(dtype))).to(dtype)
    if bias is not None:
        y = (_f(y) + _f(bias.to(dtype))).to(dtype)
    return y


def rms_norm(x: torch.Tensor, scale: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """flax nnx.RMSNorm (network.py:65-72 -> nnx.RMSNorm; contract #2).

    fp32 stats, y = x32 * (rsqrt(mean(x32^2)+eps) * scale32), one cast to x.dtype.
    """
    xf = _f(x)
    ms = (xf * xf).mean(dim=-1, keepdim=True)
    mul = torch.rsqrt(ms + eps) * _f(scale)
    return (xf * mul).to(x.dtype)


# --------------------------------------------------------------------------- RoPE
def apply_rope(inputs: torch.Tensor, positions: torch.Tensor, head_dim: int,
               max_wavelength: float = 10_000.0) -> torch.Tensor:
    """positional_embeddings.py:7-27.  inputs (B,T,N,D), positions (B|1,T) int."""
    half = head_dim // 2
    fraction = 2.0 * torch.arange(0, half, dtype=F32) / head_dim
    timescale = torch.tensor(max_wavelength, dtype=F32) ** fraction
    sinusoid = positions.to(F32)[..., None] / timescale[None, None, :]
    sinusoid = sinusoid[..., None, :]
    sin, cos = torch.sin(sinusoid), torch.cos(sinusoid)
    xf = _f(inputs)
    first, second = xf[..., :half], xf[..., half:]
    out = torch.cat([first * cos - second * sin, second * cos + first * sin], dim=-1)
    return out.to(inputs.dtype)


# --------------------------------------------------------------------------- attention
class KVCache(NamedTuple):
    key: torch.Tensor      # (B,S,Nkv,D)
    value: torch.Tensor


def dot_product_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *,
                          is_causal: bool = False,
                          kv_len: Optional[int] = None) -> torch.Tensor:
    """jax.nn.dot_product_attention, implementation="xla" (call sites attention.py:144,163;
    contract #4).  q (B,T,N,D); k,v (B,S,K,D); GQA: query head n -> kv head n // (N/K)."""
    B, T, N, D = q.shape
    S, K = k.shape[1], k.shape[2]
    G = N // K
    qf = _f(q).reshape(B, T, K, G, D)
    logits = torch.einsum("btkgd,bskd->bkgts", qf, _f(k))       # fp32 accumulate/output
    logits = logits * torch.tensor(1.0 / math.sqrt(D), dtype=F32)
    big_neg = -0.7 * torch.finfo(F32).max
    if is_causal:
        mask = torch.arange(S)[None, :] <= torch.arange(T)[:, None]
        logits = torch.where(mask[None, None, None], logits, torch.tensor(big_neg))
    if kv_len is not None:
        mask = torch.arange(S) < kv_len
        logits = torch.where(mask[None, None, None, None, :], logits, torch.tensor(big_neg))
    probs = torch.softmax(logits, dim=-1).to(k.dtype)           # probs cast to bf16 before PV
    out = torch.einsum("bkgts,bskd->btkgd", _f(probs), _f(v))
    return out.reshape(B, T, N, D).to(q.dtype)


def attention_block(p: dict, prefix: str, cfg, inputs: torch.Tensor, seq_pos: torch.Tensor,
                    kv_cache: Optional[KVCache] = None, dtype=BF16):
    """AttentionBlock.__call__ (attention.py:116-173) incl. update_kv_cache (:108-114)."""
    B, T, H = inputs.shape
    Nq, Nkv, D, S = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.max_seq_length
    wq = p[prefix + "query_proj.kernel"].reshape(H, Nq * D)
    wk = p[prefix + "key_proj.kernel"].reshape(H, Nkv * D)
    wv = p[prefix + "value_proj.kernel"].reshape(H, Nkv * D)
    wo = p[prefix + "out.kernel"].reshape(Nq * D, H)
    key = linear(inputs, wk, dtype=dtype).reshape(B, T, Nkv, D)
    value = linear(inputs, wv, dtype=dtype).reshape(B, T, Nkv, D)
    query = linear(inputs, wq, dtype=dtype).reshape(B, T, Nq, D)
    query = apply_rope(query, seq_pos, D, cfg.rope_max_wavelength)
    key = apply_rope(key, seq_pos, D, cfg.rope_max_wavelength)
    if kv_cache is not None:
        pos = int(seq_pos[0, 0]) % S                              # attention.py:109 (one scalar)
        ck, cv = kv_cache.key.clone(), kv_cache.value.clone()
        ck[:, pos:pos + T] = key
        cv[:, pos:pos + T] = value
        kv_cache = KVCache(ck, cv)
        kv_len = min(int(seq_pos[0, 0]) + 1, S)                   # attention.py:141-143
        x = dot_product_attention(query, ck, cv, kv_len=kv_len)
    else:
        assert S >= T, "sliding-window branch (attention.py:152-161) is unreachable from train_run"
        x = dot_product_attention(query, key, value, is_causal=True)
    out = linear(x.reshape(B, T, Nq * D), wo, dtype=dtype)
    return out, kv_cache


def initialize_carry(cfg, batch_size: int, dtype=BF16):
    """TransformerActorCritic.initialize_carry (network.py:297-298, attention.py:102-106)."""
    shape = (batch_size, cfg.max_seq_length, cfg.num_kv_heads, cfg.head_dim)
    return tuple(KVCache(torch.zeros(shape, dtype=dtype), torch.zeros(shape, dtype=dtype))
                 for _ in range(cfg.num_layers))


# --------------------------------------------------------------------------- GLU / block
def glu_block(p: dict, prefix: str, x: torch.Tensor, dtype=BF16) -> torch.Tensor:
    """GLUBlock.__call__ (feed_forward.py:73-78): down(gelu(up(x)) * gate(x)) (contract #5)."""
    up = linear(x, p[prefix + "up_proj.kernel"], dtype=dtype)
    gate = linear(x, p[prefix + "up_gate.kernel"], dtype=dtype)
    h = (_f(gelu_tanh(up)) * _f(gate)).to(up.dtype)
    return linear(h, p[prefix + "down_proj.kernel"], dtype=dtype)


def transformer_block(p: dict, i: int, cfg, x: torch.Tensor, time_steps: torch.Tensor,
                      carry: Optional[KVCache] = None, dtype=BF16):
    """TransformerBlock.__call__ (network.py:158-172), post-norms off (all shipped configs)."""
    pre = f"layers.{i}."
    h_in = rms_norm(x, p[pre + "history_norm.scale"])
    h_out, carry = attention_block(p, pre + "history.", cfg, h_in, time_steps, carry, dtype)
    x = (_f(x) + _f(h_out)).to(x.dtype)
    f_in = rms_norm(x, p[pre + "ffn_norm.scale"])
    f_out = glu_block(p, pre + "ffn.", f_in, dtype)
    x = (_f(x) + _f(f_out)).to(x.dtype)
    return x, carry


# --------------------------------------------------------------------------- embeds / obs encoders
def embed_encode(table: torch.Tensor, idx: torch.Tensor, dtype=BF16) -> torch.Tensor:
    """Embedder.encode (network.py:195-200): jnp.take(mode="fill", fill_value=0) wraps negative
    indices once (idx+V) and zero-fills anything still out of range; then bf16(row)*bf16(sqrt(H))."""
    V, H = table.shape
    idx = idx.long()
    idx = torch.where(idx < 0, idx + V, idx)
    valid = (idx >= 0) & (idx < V)
    rows = table[idx.clamp(0, V - 1)] * valid[..., None].to(table.dtype)
    x = rows.to(dtype)
    s = torch.tensor(math.sqrt(H), dtype=F32).to(dtype)
    return (_f(x) * _f(s)).to(dtype)


def concat_one_hot(obs: torch.Tensor, sizes, dtype=BF16) -> torch.Tensor:
    """mapox.concat_one_hot (third party; call site observation.py:85-87): per-component one-hot,
    concatenated on the last axis."""
    outs = []
    for k, n in enumerate(sizes):
        outs.append(torch.nn.functional.one_hot(obs[..., k].long().clamp(0, n - 1), n)
                    * ((obs[..., k] >= 0) & (obs[..., k] < n))[..., None])
    return torch.cat(outs, dim=-1).to(dtype)


def conv_valid(x: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor, stride, dtype=BF16):
    """nnx.Conv, padding VALID, NHWC x HWIO (observation.py:66-80)."""
    lead = x.shape[:-3]
    xf = _f(x.to(dtype)).reshape(-1, *x.shape[-3:]).permute(0, 3, 1, 2)
    wf = _f(kernel.to(dtype)).permute(3, 2, 0, 1)
    y = torch.nn.functional.conv2d(xf, wf, stride=tuple(stride)).permute(0, 2, 3, 1)
    y = y.to(dtype)
    y = (_f(y) + _f(bias.to(dtype))).to(dtype)
    return y.reshape(*lead, *y.shape[1:])


def obs_encoder(p: dict, cfg, obs: torch.Tensor, dtype=BF16) -> torch.Tensor:
    """GridCnnObsEncoder.__call__ (observation.py:84-96) / FlattenedObsEncoder (:139-145)."""
    if cfg.obs_type == "grid_cnn":
        x = concat_one_hot(obs, cfg.obs_max_value, dtype)
        n = len(cfg.cnn_kernels)
        for j in range(n):
            x = conv_valid(x, p[f"obs_encoder.layers.{j}.kernel"], p[f"obs_encoder.layers.{j}.bias"],
                           cfg.cnn_strides[j], dtype)
            if j < n - 1:
                x = gelu_tanh(x)
        return x.reshape(*x.shape[:-3], -1)
    elif cfg.obs_type == "grid_flattened":
        nc = sum(cfg.obs_max_value)
        x = torch.nn.functional.one_hot(obs.long(), nc).to(F32)          # params_dtype one-hot
        x = linear(x, p["obs_encoder.embedding.kernel"], p["obs_encoder.embedding.bias"], dtype)
        x = x.reshape(*x.shape[:-3], -1)
        return linear(x, p["obs_encoder.dense.kernel"], p["obs_encoder.dense.bias"], dtype)
    raise ValueError(cfg.obs_type)


# --------------------------------------------------------------------------- HL-Gauss
def calculate_supports(vmin: float, vmax: float, n_logits: int):
    """values.py:12-19."""
    support = torch.linspace(vmin, vmax, n_logits + 1, dtype=F32)
    centers = (support[:-1] + support[1:]) / 2
    return support[None, :], centers


def hlgauss_get_value(logits: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """HlGaussValue.get_value (values.py:43-45)."""
    return (torch.softmax(_f(logits), dim=-1) * centers).sum(-1)


def _ndtr(x: torch.Tensor) -> torch.Tensor:
    """jax.scipy.special.ndtr as used by jax.scipy.stats.norm.cdf (values.py:53):
    0.5*erfc(-x/sqrt2) style piecewise in fp32."""
    w = x * np.float32(math.sqrt(0.5))
    z = w.abs()
    y = torch.where(z < np.float32(math.sqrt(0.5)),
                    1.0 + torch.erf(w),
                    torch.where(w > 0, 2.0 - torch.erfc(z), torch.erfc(z)))
    return 0.5 * y


def hlgauss_target_probs(targets: torch.Tensor, vmin, vmax, n_logits, sigma) -> torch.Tensor:
    supports, _ = calculate_supports(vmin, vmax, n_logits)
    t = targets.reshape(-1).to(F32).clamp(vmin, vmax)
    cdf = _ndtr((supports - t[:, None]) / np.float32(sigma))
    z = cdf[:, -1] - cdf[:, 0]
    return (cdf[:, 1:] - cdf[:, :-1]) / z[:, None]


def hlgauss_get_loss(logits: torch.Tensor, targets: torch.Tensor, vmin, vmax, sigma) -> torch.Tensor:
    """HlGaussValue.get_loss (values.py:47-63): soft-label CE, mean over tokens."""
    n_logits = logits.shape[-1]
    lg = _f(logits).reshape(-1, n_logits)
    probs = hlgauss_target_probs(targets, vmin, vmax, n_logits, sigma)
    return -(probs * torch.log_softmax(lg, dim=-1)).sum(-1).mean()


# --------------------------------------------------------------------------- trunk
class TimeStep(NamedTuple):
    """mapox.TimeStep (third party; fields per SURVEY appendix A)."""
    obs: torch.Tensor
    time: torch.Tensor
    terminated: torch.Tensor
    last_action: torch.Tensor
    reward: torch.Tensor
    action_mask: Optional[torch.Tensor] = None
    task_ids: Optional[torch.Tensor] = None


def model_embed(p: dict, cfg, ts: TimeStep, dtype=BF16) -> torch.Tensor:
    """network.py:303-309."""
    x = obs_encoder(p, cfg, ts.obs, dtype)
    r = linear(ts.reward[..., None].to(F32), p["reward_encoder.kernel"], p["reward_encoder.bias"], dtype)
    a = embed_encode(p["action_embedder.embedding_table"], ts.last_action, dtype)
    x = (_f(x) + _f(r)).to(dtype)
    x = (_f(x) + _f(a)).to(dtype)
    if ts.task_ids is not None and cfg.task_count > 1:
        x = (_f(x) + _f(embed_encode(p["task_embedder.embedding_table"], ts.task_ids, dtype))).to(dtype)
    return x


def model_heads(p: dict, cfg, x: torch.Tensor, action_mask: Optional[torch.Tensor], dtype=BF16):
    """network.py:321-341: out norm, tied fp32 logits + mask, value MLP + gelu, fp32 value head."""
    x = rms_norm(x, p["output_norm.scale"])
    logits = _f(x) @ _f(p["action_embedder.embedding_table"]).T
    if action_mask is not None:
        logits = torch.where(action_mask.bool(), logits, torch.tensor(torch.finfo(F32).min))
    prevalue = x
    if cfg.value_hidden_dim is not None:
        prevalue = gelu_tanh(linear(prevalue, p["value_mlp.kernel"], p["value_mlp.bias"], dtype))
    vlogits = linear(prevalue, p["value_head.dense.kernel"], p["value_head.dense.bias"], dtype=None)
    return vlogits, logits


def model_forward(p: dict, cfg, ts: TimeStep, carry=None, dtype=BF16, return_hidden: bool = False):
    """TransformerActorCritic.__call__ (network.py:300-343).
    Returns (value_rep f32 (B,T,NL), action_logits f32 (B,T,A), carry)."""
    x = model_embed(p, cfg, ts, dtype)
    out_carry = []
    for i in range(cfg.num_layers):
        x, c = transformer_block(p, i, cfg, x, ts.time, None if carry is None else carry[i], dtype)
        out_carry.append(c)
    hidden = x
    vlogits, logits = model_heads(p, cfg, x, ts.action_mask, dtype)
    if return_hidden:
        return vlogits, logits, (tuple(out_carry) if carry is not None else None), hidden
    return vlogits, logits, (tuple(out_carry) if carry is not None else None)


# --------------------------------------------------------------------------- categorical (distrax)
def categorical_log_prob(logits: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    return torch.log_softmax(logits, dim=-1).gather(-1, actions.long()[..., None]).squeeze(-1)


def categorical_entropy(logits: torch.Tensor) -> torch.Tensor:
    lp = torch.log_softmax(logits, dim=-1)
    pl = torch.where(torch.isinf(lp), torch.zeros_like(lp), lp * lp.exp())
    return -pl.sum(-1)


def categorical_sample_gumbel(logits: torch.Tensor, gumbel: torch.Tensor) -> torch.Tensor:
    """distrax.Categorical.sample == argmax(logits + Gumbel noise) (train.py:97); the noise is an
    input so the index is a pure function of (logits, noise).  First max wins (jnp.argmax)."""
    return torch.argmax(logits + gumbel, dim=-1).to(torch.int32)


# --------------------------------------------------------------------------- GAE
def calculate_advantage(rewards: np.ndarray, values: np.ndarray, next_terminated: np.ndarray,
                        discount: float, gae_lambda: float, norm_adv: bool):
    """Rollout.calculate_advantage (rollout.py:128-167) in numpy fp32, sequential in t,
    one fp32 rounding per arithmetic op (no fused multiply-add)."""
    r = np.asarray(rewards, np.float32)
    v = np.asarray(values, np.float32)
    term = np.asarray(next_terminated).astype(bool)
    B, T = r.shape
    lam = np.float32(gae_lambda)
    one_m = np.float32(1.0 - gae_lambda)         # python-float weak type folded before the scan
    d = np.where(term, np.float32(0.0), np.float32(discount)).astype(np.float32)
    acc = v[:, T].copy()
    targets = np.empty((B, T), np.float32)
    for t in range(T - 1, -1, -1):
        acc = r[:, t] + d[:, t] * (one_m * v[:, t + 1] + lam * acc)
        targets[:, t] = acc
    adv = targets - v[:, :T]
    if norm_adv:
        mean = adv.mean(dtype=np.float32)
        std = np.sqrt(((adv - mean) ** 2).mean(dtype=np.float32), dtype=np.float32)
        adv = (adv - mean) / (std + np.float32(1e-8))
    return adv.astype(np.float32), targets


# --------------------------------------------------------------------------- PPO loss
def ppo_loss(p: dict, cfg, hyp, batch: dict, progress: float, dtype=BF16):
    """ppo_loss (train.py:162-222).  ``batch`` holds the RolloutState minibatch fields (B,T,...)."""
    T = batch["obs"].shape[1]
    positions = torch.arange(T, dtype=torch.int32)[None, :]
    ts = TimeStep(obs=batch["obs"], time=positions, terminated=batch["terminated"],
                  last_action=batch["last_actions"], reward=batch["last_rewards"],
                  action_mask=batch["action_mask"], task_ids=batch.get("task_ids"))
    vlogits, logits, _ = model_forward(p, cfg, ts, None, dtype)
    log_probs = categorical_log_prob(logits, batch["actions"])
    value_loss = hlgauss_get_loss(vlogits, batch["targets"], cfg.value_min, cfg.value_max, cfg.value_sigma)
    ratio = torch.exp(log_probs - batch["log_prob"])
    adv = batch["advantages"]
    pg1 = ratio * adv
    pg2 = torch.clamp(ratio, 1.0 - hyp.vf_clip, 1.0 + hyp.vf_clip) * adv     # vf_clip is the policy eps (train.py:201)
    actor_loss = -torch.minimum(pg1, pg2).mean()
    entropy_loss = -categorical_entropy(logits).mean()
    ent_coef = hyp.entropy_coef + progress * (hyp.entropy_coef_end - hyp.entropy_coef)
    total = hyp.vf_coef * value_loss + actor_loss + ent_coef * entropy_loss
    return total, dict(value_loss=value_loss, actor_loss=actor_loss, entropy_loss=entropy_loss,
                       total_loss=total, entropy_coef=ent_coef)


# --------------------------------------------------------------------------- rollout (evaluate)
def rollout_decode(p: dict, cfg, obs_seq: torch.Tensor, reward_seq: torch.Tensor,
                   gumbel_seq: torch.Tensor, action_mask_seq: Optional[torch.Tensor] = None,
                   task_ids: Optional[torch.Tensor] = None, dtype=BF16):
    """evaluate()'s fori_loop (train.py:78-140) against a synthetic environment whose next
    observation / reward are pre-drawn (north star: "synthetic observations") and which echoes the
    sampled action as next_timestep.last_action.  obs_seq (T+1,B,...), reward_seq (T+1,B) (reward
    delivered WITH obs t, i.e. timestep.reward), gumbel_seq (T,B,A).
    Returns dict of RolloutState fields (B,T,...) with values (B,T+1) (bootstrap per train.py:142-148
    quirk: value of the initial timestep with a zero cache)."""
    Tn = obs_seq.shape[0] - 1
    B = obs_seq.shape[1]
    centers = calculate_supports(cfg.value_min, cfg.value_max, cfg.value_n_logits)[1]
    carry = initialize_carry(cfg, B, dtype)
    last_action = torch.zeros(B, dtype=torch.int32)
    out = dict(actions=[], log_prob=[], values=[], last_actions=[], last_rewards=[], rewards=[], logits=[])

    def mk(t, la):
        am = None if action_mask_seq is None else action_mask_seq[t][:, None]
        return TimeStep(obs=obs_seq[t][:, None], time=torch.full((B, 1), t, dtype=torch.int32),
                        terminated=torch.zeros(B, 1, dtype=torch.bool), last_action=la[:, None],
                        reward=reward_seq[t][:, None], action_mask=am,
                        task_ids=None if task_ids is None else task_ids[:, None])

    for t in range(Tn):
        ts = mk(t, last_action)
        vlogits, logits, carry = model_forward(p, cfg, ts, carry, dtype)
        action = categorical_sample_gumbel(logits[:, 0], gumbel_seq[t])
        lp = categorical_log_prob(logits[:, 0], action)
        value = hlgauss_get_value(vlogits[:, 0], centers)
        out["actions"].append(action); out["log_prob"].append(lp); out["values"].append(value)
        out["logits"].append(logits[:, 0])
        out["last_actions"].append(last_action.clone()); out["last_rewards"].append(reward_seq[t])
        out["rewards"].append(reward_seq[t + 1])
        last_action = action
    # bootstrap: initial timestep, initial zero carry (train.py:142-148)
    ts0 = mk(0, torch.zeros(B, dtype=torch.int32))
    vlogits, _, _ = model_forward(p, cfg, ts0, initialize_carry(cfg, B, dtype), dtype)
    out["values"].append(hlgauss_get_value(vlogits[:, 0], centers))
    res = {k: torch.stack(v, dim=1) for k, v in out.items()}
    res["carry"] = carry
    return res


# --------------------------------------------------------------------------- optimizer (parity unpinned)
def _ns_orthogonalize(x: torch.Tensor, steps: int = 5, eps: float = 1e-8, coeffs=(3.4445, -4.7750, 2.0315)):
    """optax.contrib orthogonalize_via_newton_schulz for a 2-D matrix (fp32)."""
    transposed = x.shape[0] > x.shape[1]
    if transposed:
        x = x.t()
    x = x / (torch.linalg.norm(x) + eps)
    c0, c1, c2 = coeffs
    for _ in range(steps):
        a = x @ x.t()
        b = c1 * a + c2 * (a @ a)
        x = c0 * x + b @ x
    return x.t() if transposed else x


def optimizer_step(params: dict, grads: dict, state: dict, t: int, cfg, total_steps: int, is_muon, muon_beta=0.95):
    """One step of create_optimizer (optimizer.py:7-39) restated from optax 0.2.8: muon partition (nesterov momentum,
    Newton-Schulz, sqrt(max(1,out/in)), decayed weights, linear LR) + adamw(nesterov=True) for the rest when
    cfg.type == "muon"; plain optax.adamw when "adamw"; then clip_by_global_norm on the UPDATE.  Returns new params.
    state: {"mu": {...}, "nu": {...}} (zeros initially); t is the 0-based step count."""
    lr = cfg.learning_rate * (1.0 - min(t, total_steps) / total_steps)
    upd = {}
    for k, p in params.items():
        g = grads[k]
        if cfg.type == "muon" and is_muon(k, tuple(p.shape)):
            m = muon_beta * state["mu"][k] + (1 - muon_beta) * g
            state["mu"][k] = m
            mh = muon_beta * (m / (1 - muon_beta ** (t + 2))) + (1 - muon_beta) * (g / (1 - muon_beta ** (t + 1)))
            o = _ns_orthogonalize(mh)
            o = o * math.sqrt(max(1.0, p.shape[1] / p.shape[0]))
            upd[k] = -lr * (o + cfg.weight_decay * p)
        else:
            b1, b2 = cfg.beta1, cfg.beta2
            eps = 1e-8 if cfg.type == "muon" else cfg.eps
            m = b1 * state["mu"][k] + (1 - b1) * g
            v = b2 * state["nu"][k] + (1 - b2) * g * g
            state["mu"][k], state["nu"][k] = m, v
            if cfg.type == "muon":
                mh = b1 * (m / (1 - b1 ** (t + 2))) + (1 - b1) * (g / (1 - b1 ** (t + 1)))
            else:
                mh = m / (1 - b1 ** (t + 1))
            vh = v / (1 - b2 ** (t + 1))
            upd[k] = -lr * (mh / (torch.sqrt(vh) + eps) + cfg.weight_decay * p)
    gn = math.sqrt(sum(float((u.double() ** 2).sum()) for u in upd.values()))
    scale = cfg.max_norm / gn if gn >= cfg.max_norm else 1.0
    return {k: params[k] + upd[k] * scale for k in params}, gn
