"""Plain-struct view of the reference's config tree (mapox_trainer/config.py) - only the fields the hot path reads.

``from_reference_json`` accepts the reference's own JSON files (config/*.json); the env facts that live in
mapox (observation components, action count) are parameters because they are not in the JSON.
``NAMED`` holds the shapes of the five configs BASELINE.json names.
"""

from __future__ import annotations

import json
from dataclasses import dataclass, field, replace
from typing import Optional, Tuple


@dataclass(frozen=True)
class ModelConfig:
    # TransformerActorCriticConfig (config.py:80-101) + env facts
    hidden_features: int = 128
    num_layers: int = 2
    num_heads: int = 4
    num_kv_heads: int = 1
    head_dim: int = 32
    rope_max_wavelength: float = 10_000.0
    ffn_size: int = 768
    value_hidden_dim: Optional[int] = 768
    value_min: float = -10.0
    value_max: float = 10.0
    value_n_logits: int = 51
    value_sigma: float = 0.10381270166884848
    obs_type: str = "grid_cnn"
    cnn_kernels: Tuple[Tuple[int, int], ...] = ((3, 3), (3, 3), (3, 3))
    cnn_strides: Tuple[Tuple[int, int], ...] = ((2, 2), (1, 1), (1, 1))
    cnn_channels: Tuple[int, ...] = (16, 32)
    obs_shape: Tuple[int, ...] = (11, 11, 2)           # (view_w, view_h, K) grid_cnn; (view_w, view_h) grid_flattened
    obs_max_value: Tuple[int, ...] = (5, 5)            # per-component one-hot sizes (ObservationSpec.max_value)
    action_dim: int = 8                                # ActionSpec.n
    task_count: int = 1
    max_seq_length: int = 512                          # = max_env_steps (train.py:353)

    @property
    def q_dim(self) -> int:
        return self.num_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.num_kv_heads * self.head_dim


@dataclass(frozen=True)
class PPOConfig:
    # PPOConfig (config.py:124-138)
    epoch_count: int = 1
    minibatch_count: int = 16
    vf_coef: float = 0.5392535826611433
    entropy_coef: float = 0.001
    entropy_coef_end: float = 0.0
    vf_clip: float = 0.2827893540340816
    discount: float = 0.9660902557298033
    gae_lambda: float = 0.9123716252262187
    normalize_advantage: bool = True


@dataclass(frozen=True)
class OptimizerConfig:
    # OptimizerConfig (config.py:112-121)
    type: str = "muon"
    learning_rate: float = 0.001
    weight_decay: float = 0.00010718901435954378
    eps: float = 1e-8
    beta1: float = 0.9353521064414686
    beta2: float = 0.9040155258546868
    max_norm: float = 0.39644292421090327


@dataclass(frozen=True)
class RunConfig:
    name: str
    model: ModelConfig
    ppo: PPOConfig = field(default_factory=PPOConfig)
    optimizer: OptimizerConfig = field(default_factory=OptimizerConfig)
    num_agents: int = 4096        # B = num_envs x agents per env
    update_steps: int = 5000
    seed: int = 42

    @property
    def optimizer_total_steps(self) -> int:
        """Horizon of the linear LR schedule: the reference passes update_steps * minibatch_count * epoch_count to
        create_optimizer (train.py:360-365) and calls optimizer.update once per minibatch (train.py:244), so the
        schedule reaches 0 exactly at the end of training."""
        return self.update_steps * self.ppo.minibatch_count * self.ppo.epoch_count


def from_reference_json(path: str, *, obs_components: Tuple[int, ...] = (5, 5), action_dim: int = 8,
                        agents_per_env: Optional[int] = None, task_count: int = 1) -> RunConfig:
    """Read a reference config/*.json (mapox_trainer/config.py:167-172 load_config) into a RunConfig."""
    with open(path) as f:
        j = json.load(f)
    m = j["learner"]["model"]
    hist = m["layer"]["history"]
    enc = m["obs_encoder"]
    env = j["environment"]
    val = m["value"]
    unsupported = _unsupported_model_options(m)
    if unsupported:
        raise ValueError("reference config options outside the B200 hot path (the engine would run a different "
                         "network): " + "; ".join(unsupported))
    model = ModelConfig(
        hidden_features=m["hidden_features"], num_layers=m["num_layers"], num_heads=hist["num_heads"],
        num_kv_heads=hist["num_kv_heads"], head_dim=hist["head_dim"],
        rope_max_wavelength=hist.get("rope_max_wavelength", 10_000.0), ffn_size=m["layer"]["feed_forward"]["size"],
        value_hidden_dim=m.get("value_hidden_dim"), value_min=val["min"], value_max=val["max"],
        value_n_logits=val["n_logits"], value_sigma=val["sigma"], obs_type=enc["obs_type"],
        cnn_kernels=tuple(tuple(k) for k in enc.get("kernels", ())),
        cnn_strides=tuple(tuple(s) for s in enc.get("strides", ())), cnn_channels=tuple(enc.get("channels", ())),
        obs_shape=(env.get("view_width", 11), env.get("view_height", 11), len(obs_components)),
        obs_max_value=tuple(obs_components), action_dim=action_dim, task_count=task_count,
        max_seq_length=j["max_env_steps"])
    t = j["learner"]["trainer"]
    ppo = PPOConfig(epoch_count=t.get("epoch_count", 1), minibatch_count=t.get("minibatch_count", 1),
                    vf_coef=t.get("vf_coef", 0.05), entropy_coef=t.get("entropy_coef", 0.005),
                    entropy_coef_end=t.get("entropy_coef_end", 0.0) or 0.0, vf_clip=t.get("vf_clip", 0.2),
                    discount=t.get("discount", 0.95), gae_lambda=t.get("gae_lambda", 0.95),
                    normalize_advantage=t.get("normalize_advantage", False))
    o = j["learner"]["optimizer"]
    opt = OptimizerConfig(type=o["type"], learning_rate=o["learning_rate"], weight_decay=o["weight_decay"],
                          eps=o["eps"], beta1=o["beta1"], beta2=o["beta2"], max_norm=o["max_norm"])
    if env.get("env_type") == "multi":
        # MultiTaskConfig (mapox, third party): one task id per sub-environment, agents = sum(num * agents per env);
        # envs whose per-env agent count lives in mapox (scouts) take `agents_per_env`
        def _agents(e):
            if "num_agents" in e:
                return e["num_agents"]
            if "num_sneakers" in e:
                return e["num_sneakers"] + e.get("num_chasers", 0)
            if agents_per_env is None:
                raise ValueError(f"agents per env of {e.get('env_type')!r} lives in mapox: pass agents_per_env")
            return agents_per_env
        total_agents = sum(e["num"] * _agents(e["env"]) for e in env["envs"])
        model = replace(model, task_count=len(env["envs"]))
    else:
        apn = agents_per_env if agents_per_env is not None else env.get("num_agents", 1)
        total_agents = j.get("num_envs", 1) * apn
    return RunConfig(name=path, model=model, ppo=ppo, optimizer=opt, num_agents=total_agents,
                     update_steps=j["update_steps"], seed=j["seed"] if isinstance(j["seed"], int) else 42)


def _unsupported_model_options(m: dict) -> list:
    """The engine hardcodes what every config BASELINE.json names uses: tanh-GELU GLU blocks,
    RMSNorm, attention history without qk-norm / post-norms, bf16 compute over fp32 parameters, an HL-Gauss head.
    Anything else must be rejected, not silently run as a different network (TransformerActorCriticConfig,
    mapox_trainer/config.py:80-101 with its defaults)."""
    layer = m.get("layer", {})
    hist = layer.get("history") or {}
    bad = []
    if m.get("value", {}).get("type") != "hl_gauss":
        bad.append("value.type != hl_gauss (mse is craftax-only)")
    if m.get("activation") != "gelu":
        bad.append(f"activation={m.get('activation')!r} (gelu only)")
    if m.get("norm") != "rms_norm":
        bad.append(f"norm={m.get('norm')!r} (rms_norm only)")
    if m.get("dtype", "float32") != "bfloat16":
        bad.append(f"dtype={m.get('dtype', 'float32')!r} (bfloat16 only)")
    if m.get("param_dtype", "float32") != "float32":
        bad.append(f"param_dtype={m.get('param_dtype')!r} (float32 only)")
    if m.get("kernel_init", "glorot_uniform") != "glorot_uniform":
        bad.append(f"kernel_init={m.get('kernel_init')!r} (glorot_uniform only; affects initialisation, not kernels)")
    if hist.get("type", "attention") != "attention" or not hist:
        bad.append(f"layer.history.type={hist.get('type')!r} (attention only)")
    if hist.get("use_qk_norm", False):
        bad.append("layer.history.use_qk_norm=true")
    if layer.get("use_post_attn_norm", False) or layer.get("use_post_ffw_norm", False):
        bad.append("layer.use_post_attn_norm / use_post_ffw_norm")
    if not layer.get("feed_forward", {}).get("glu", True):
        bad.append("layer.feed_forward.glu=false")
    if m.get("obs_encoder", {}).get("obs_type") not in ("grid_cnn", "grid_flattened"):
        bad.append(f"obs_encoder.obs_type={m.get('obs_encoder', {}).get('obs_type')!r}")
    return bad


_BASE = ModelConfig()
NAMED = {
    # tiny find_return-shaped plumbing config (config/test.json is stale)
    "test": RunConfig("test", replace(_BASE, max_seq_length=64, obs_shape=(11, 11, 2)), num_agents=16,
                      ppo=PPOConfig(minibatch_count=2)),
    "return_2_layers": RunConfig("return_2_layers", _BASE),
    "return_8_layers": RunConfig("return_8_layers", replace(_BASE, num_layers=8)),
    "multitask": RunConfig("multitask", replace(_BASE, hidden_features=256, num_layers=16, task_count=4),
                           ppo=PPOConfig(minibatch_count=32, discount=0.99, gae_lambda=0.91),
                           optimizer=OptimizerConfig(learning_rate=0.003)),
    "blog_32_layers": RunConfig("blog_32_layers",
                                replace(_BASE, num_layers=32, obs_type="grid_flattened", obs_shape=(5, 5),
                                        obs_max_value=(8,)),
                                ppo=PPOConfig(normalize_advantage=False, entropy_coef=8.270466882977551e-05),
                                update_steps=3800),
}
