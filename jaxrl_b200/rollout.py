"""Mirror of mapox_trainer/rollout.py: RolloutState / Rollout with calculate_advantage on the B200 GAE kernel."""

from __future__ import annotations

from types import SimpleNamespace
from typing import NamedTuple

import torch

from . import ops

F32, I32 = torch.float32, torch.int32


def ObservationSpec(shape, dtype=torch.uint8, max_value=None):
    return SimpleNamespace(shape=tuple(shape), dtype=dtype, max_value=max_value)


def ActionSpec(n: int):
    return SimpleNamespace(n=n)


class RolloutState(NamedTuple):     # rollout.py:11-30
    obs: torch.Tensor
    action_mask: torch.Tensor
    actions: torch.Tensor
    rewards: torch.Tensor
    terminated: torch.Tensor
    next_terminated: torch.Tensor
    log_prob: torch.Tensor
    values: torch.Tensor
    advantages: torch.Tensor
    targets: torch.Tensor
    last_actions: torch.Tensor
    last_rewards: torch.Tensor
    task_ids: torch.Tensor


class Rollout:
    def __init__(self, batch_size: int, trajectory_length: int, obs_spec, action_spec, device="cuda"):
        self.batch_size, self.trajectory_length = batch_size, trajectory_length
        self.obs_spec, self.action_spec = obs_spec, action_spec
        self.device = torch.device(device)

    def create_state(self) -> RolloutState:
        B, T, dev = self.batch_size, self.trajectory_length, self.device
        z = lambda *s, dtype=F32: torch.zeros(*s, dtype=dtype, device=dev)
        return RolloutState(
            obs=z(B, T, *self.obs_spec.shape, dtype=self.obs_spec.dtype),
            action_mask=torch.ones(B, T, self.action_spec.n, dtype=torch.bool, device=dev),
            actions=z(B, T, dtype=I32), rewards=z(B, T), terminated=z(B, T, dtype=torch.bool),
            next_terminated=z(B, T, dtype=torch.bool), log_prob=z(B, T), values=z(B, T + 1), advantages=z(B, T),
            targets=z(B, T), last_actions=z(B, T, dtype=I32), last_rewards=z(B, T), task_ids=z(B, T, dtype=I32))

    def store(self, state: RolloutState, step: int, timestep, next_timestep, log_prob, value) -> RolloutState:
        """rollout.py:93-126 (in place on the state's buffers)."""
        state.obs[:, step] = timestep.obs
        if timestep.action_mask is not None:
            state.action_mask[:, step] = timestep.action_mask
        state.actions[:, step] = next_timestep.last_action
        state.log_prob[:, step] = log_prob
        state.values[:, step] = value
        state.rewards[:, step] = next_timestep.reward
        state.terminated[:, step] = timestep.terminated
        state.next_terminated[:, step] = next_timestep.terminated
        state.last_actions[:, step] = timestep.last_action
        state.last_rewards[:, step] = timestep.reward
        if timestep.task_ids is not None:
            state.task_ids[:, step] = timestep.task_ids
        return state

    def calculate_advantage(self, state: RolloutState, discount: float, gae_lambda: float, norm_adv: bool) -> RolloutState:
        """rollout.py:128-167."""
        adv, tgt, stats = ops.gae(state.rewards.to(F32).contiguous(), state.values.to(F32).contiguous(),
                                  state.next_terminated.contiguous(), discount, gae_lambda)
        if norm_adv:
            ops.adv_normalize(adv.view(-1), stats)
        return state._replace(advantages=adv, targets=tgt)

    def create_minibatches(self, state: RolloutState, minibatches: int, rng: torch.Generator = None) -> RolloutState:
        """rollout.py:169-181: shuffle rows, then '(m b) ... -> m b ...'."""
        if minibatches > 1:
            idx = torch.randperm(self.batch_size, generator=rng, device=self.device if rng is None or rng.device.type == "cuda" else "cpu").to(self.device)
            state = RolloutState(*[x[idx] for x in state])
        return RolloutState(*[x.reshape(minibatches, self.batch_size // minibatches, *x.shape[1:]) for x in state])
