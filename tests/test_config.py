"""CPU: config reader - rejects reference options the engine does not implement, LR-schedule horizon."""

import glob
import json
import os

import pytest

from jaxrl_b200.config import NAMED, _unsupported_model_options, from_reference_json

MODEL = {
    "obs_encoder": {"obs_type": "grid_cnn", "kernels": [[3, 3]] * 3, "strides": [[2, 2], [1, 1], [1, 1]], "channels": [16, 32]},
    "hidden_features": 128,
    "layer": {"feed_forward": {"glu": True, "size": 768},
              "history": {"type": "attention", "num_heads": 4, "num_kv_heads": 1, "head_dim": 32, "use_qk_norm": False},
              "use_post_attn_norm": False, "use_post_ffw_norm": False},
    "num_layers": 2, "value_hidden_dim": 768,
    "value": {"type": "hl_gauss", "min": -10.0, "max": 10.0, "n_logits": 51, "sigma": 0.1},
    "activation": "gelu", "norm": "rms_norm", "kernel_init": "glorot_uniform", "dtype": "bfloat16", "param_dtype": "float32",
}


def _with(path, value):
    m = json.loads(json.dumps(MODEL))
    d = m
    for k in path[:-1]:
        d = d[k]
    d[path[-1]] = value
    return m


def test_supported_subset_is_accepted():
    assert _unsupported_model_options(MODEL) == []


@pytest.mark.parametrize("path,value", [
    (("activation",), "relu"), (("norm",), "layer_norm"), (("dtype",), "float32"), (("param_dtype",), "bfloat16"),
    (("layer", "history", "type"), "rnn"), (("layer", "history", "use_qk_norm"), True),
    (("layer", "use_post_attn_norm"), True), (("layer", "use_post_ffw_norm"), True),
    (("layer", "feed_forward", "glu"), False), (("value", "type"), "mse"), (("obs_encoder", "obs_type"), "linear"),
    (("kernel_init",), "he_uniform")])
def test_options_outside_the_hot_path_are_rejected(path, value):
    assert _unsupported_model_options(_with(path, value)), path


def test_from_reference_json_raises(tmp_path):
    cfg = {"seed": 1, "max_env_steps": 64, "num_envs": 2, "update_steps": 10,
           "environment": {"num_agents": 4, "view_width": 11, "view_height": 11},
           "learner": {"model": _with(("norm",), "layer_norm"),
                       "trainer": {"minibatch_count": 2}, "optimizer": {"type": "muon", "learning_rate": 1e-3,
                                                                        "weight_decay": 0.0, "eps": 1e-8, "beta1": 0.9,
                                                                        "beta2": 0.9, "max_norm": 0.5}}}
    f = tmp_path / "c.json"
    f.write_text(json.dumps(cfg))
    with pytest.raises(ValueError, match="layer_norm"):
        from_reference_json(str(f))
    cfg["learner"]["model"] = MODEL
    f.write_text(json.dumps(cfg))
    run = from_reference_json(str(f))
    assert run.num_agents == 8 and run.model.max_seq_length == 64


@pytest.mark.skipif(not os.path.isdir("/root/reference/config"), reason="reference checkout not present")
def test_named_configs_match_the_reference_files():
    """The five shapes BASELINE.json names, read from the reference's own JSON (only where the checkout exists)."""
    for name in ("return_2_layers", "return_8_layers", "blog_32_layers"):
        kw = dict(obs_components=(8,)) if name == "blog_32_layers" else {}
        run = from_reference_json(f"/root/reference/config/{name}.json", **kw)
        ours = NAMED[name]
        for f in ("hidden_features", "num_layers", "num_heads", "num_kv_heads", "head_dim", "ffn_size", "value_hidden_dim",
                  "value_n_logits", "value_sigma", "obs_type", "max_seq_length"):
            assert getattr(run.model, f) == getattr(ours.model, f), (name, f)
        assert run.ppo == ours.ppo and run.optimizer == ours.optimizer and run.num_agents == ours.num_agents
        assert run.update_steps == ours.update_steps
    mt = from_reference_json("/root/reference/config/multitask.json", agents_per_env=4)
    assert mt.model.hidden_features == 256 and mt.model.num_layers == 16 and mt.model.task_count == 4
    assert mt.ppo == NAMED["multitask"].ppo and mt.optimizer == NAMED["multitask"].optimizer
    for f in glob.glob("/root/reference/config/craftax*.json"):
        with pytest.raises(ValueError):
            from_reference_json(f)


def test_lr_schedule_horizon_follows_the_reference_call_site():
    """train.py:360-365 passes update_steps * minibatch_count * epoch_count to create_optimizer, whose
    optax.linear_schedule(lr, 0, n) is lr * (1 - min(t, n) / n): with one optimizer step per minibatch the rate reaches
    0 at the end of training, not after update_steps / minibatch_count outer updates."""
    run = NAMED["return_2_layers"]
    n = run.optimizer_total_steps
    assert n == run.update_steps * 16
    from oracle import oracle as O
    import torch
    p = {"w": torch.ones(2, 2)}
    g = {"w": torch.ones(2, 2)}
    st = {"mu": {"w": torch.zeros(2, 2)}, "nu": {"w": torch.zeros(2, 2)}}
    from dataclasses import replace
    ocfg = replace(run.optimizer, type="adamw", max_norm=1e9, weight_decay=0.0, beta1=0.0, beta2=0.0)   # update = lr * sign(g)
    half, _ = O.optimizer_step(p, g, st, n // 2, ocfg, n, lambda *_: False)
    st = {"mu": {"w": torch.zeros(2, 2)}, "nu": {"w": torch.zeros(2, 2)}}
    first, _ = O.optimizer_step(p, g, st, 0, ocfg, n, lambda *_: False)
    ratio = ((half["w"] - 1).abs().mean() / (first["w"] - 1).abs().mean()).item()
    assert abs(ratio - 0.5) < 1e-3
