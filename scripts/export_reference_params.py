#!/usr/bin/env python
"""Export the parameters of a mapox_trainer run to a flat .npz that jaxrl_b200 reads (ParamStore.load_npz).

Run this on the REFERENCE side (an environment with jax / flax / orbax / mapox, i.e. not this repo's image):

    python export_reference_params.py <run name> [--base-dir results] [--out params.npz]

It restores the latest checkpoint exactly the way the reference's own tools do (play.py:load_policy ->
Checkpointer.restore_latest, checkpointer.py:19-38: an orbax StandardRestore of nnx.state(model, nnx.Param)) and writes
one array per leaf, keyed by its nnx path joined with "." and stripped of the trailing ".value" - e.g.
"layers.0.history.query_proj.kernel" (H, Nq, D) float32 - which are the names of jaxrl_b200.params.param_specs.
With --golden N it also runs the reference model on N seeded synthetic decode steps (attention_impl="xla") and
stores inputs and outputs, so that the pair (params, golden vectors) can be committed under tests/golden/ and pin the
oracle to real reference outputs.
"""

import argparse

import numpy as np


def flatten_state(state) -> dict:
    """nnx.State -> {"a.b.0.c": np.ndarray} (paths as in jaxrl_b200.params.param_specs)."""
    out = {}
    for path, leaf in state.flat_state():
        keys = [str(k) for k in path]
        if keys and keys[-1] == "value":
            keys = keys[:-1]
        value = getattr(leaf, "value", leaf)
        out[".".join(keys)] = np.asarray(value, dtype=np.float32)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("run")
    ap.add_argument("--base-dir", default="results")
    ap.add_argument("--out", default="params.npz")
    ap.add_argument("--golden", type=int, default=0, help="also store N seeded decode steps of the reference model")
    args = ap.parse_args()

    import jax
    import jax.numpy as jnp
    from flax import nnx
    from mapox import EnvironmentFactory
    from mapox_trainer.experiment import Experiment
    from mapox_trainer.play import load_policy

    experiment = Experiment.load(args.run, args.base_dir)
    config = experiment.config
    rngs = nnx.Rngs(0)
    env, task_count = EnvironmentFactory().create_env(config.environment, config.max_env_steps, 1)
    model = load_policy(experiment, env, config.max_env_steps, task_count, rngs)
    arrays = flatten_state(nnx.state(model, nnx.Param))
    meta = dict(obs_shape=np.asarray(env.observation_spec.shape), obs_max_value=np.asarray(env.observation_spec.max_value),
                action_dim=np.asarray(env.action_spec.n), task_count=np.asarray(task_count),
                max_seq_length=np.asarray(config.max_env_steps))
    if args.golden > 0:
        from mapox import TimeStep
        rng = np.random.default_rng(0)
        B = 8
        spec = env.observation_spec
        mv = np.broadcast_to(np.asarray(spec.max_value), spec.shape[-1:]) if len(spec.shape) == 3 else np.asarray(spec.max_value)
        obs = (rng.random((args.golden, B) + tuple(spec.shape)) * mv).astype(np.uint8)
        reward = rng.standard_normal((args.golden, B)).astype(np.float32) * 0.1
        last_action = rng.integers(0, env.action_spec.n, (args.golden, B)).astype(np.int32)
        carry = model.initialize_carry(B, rngs)
        vl, lg = [], []
        for t in range(args.golden):
            ts = TimeStep(obs=jnp.asarray(obs[t])[:, None], time=jnp.full((B, 1), t, jnp.int32),
                          terminated=jnp.zeros((B, 1), bool), last_action=jnp.asarray(last_action[t])[:, None],
                          reward=jnp.asarray(reward[t])[:, None], action_mask=None, task_ids=None)
            value_rep, policy, carry = model(ts, carry)
            vl.append(np.asarray(value_rep[:, 0], np.float32))
            lg.append(np.asarray(policy.logits[:, 0], np.float32))
        meta.update(golden_obs=obs, golden_reward=reward, golden_last_action=last_action,
                    golden_value_logits=np.stack(vl), golden_action_logits=np.stack(lg),
                    golden_backend=np.asarray(jax.default_backend()))
    np.savez(args.out, **{"param/" + k: v for k, v in arrays.items()}, **{"meta/" + k: v for k, v in meta.items()})
    print(f"wrote {len(arrays)} parameter arrays to {args.out}")


if __name__ == "__main__":
    main()
