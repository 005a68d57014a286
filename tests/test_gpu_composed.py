"""Composed parity: one whole PPOTrainer.train_step() (rollout -> layout change + row permutation -> GAE ->
minibatch forward/backward -> optimizer) against the oracle pipeline on the same seeds and inputs, i.e. the
reference's train() (train.py:225-309): evaluate() (:42-159), Rollout.store semantics (rollout.py:93-126),
calculate_advantage (:128-167), create_minibatches (:169-181), ppo_loss (train.py:162-222), optimizer.update.

The oracle is fed the kernels' own Gumbel noise (mpx_gumbel_noise with the trainer's seed / stream ids) and the
trainer's own row permutation, so every stage can be compared element by element:
  * sampled actions bit-equal until an agent's first near-tie (the oracle's top-2 margin there must be tiny),
  * the batch-major training arrays (which slice of rewards / actions goes where) exactly,
  * GAE <= 1e-5, targets bit exact given the same inputs,
  * the gradient of a minibatch and the minibatch losses,
  * the parameter update of the full step.
"""

from dataclasses import replace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from jaxrl_b200.config import NAMED  # noqa: E402
from jaxrl_b200.params import is_muon_param  # noqa: E402
from oracle import oracle as O  # noqa: E402
from _util import log as _log  # noqa: E402

DEV = "cuda"
RUN = replace(NAMED["test"], num_agents=32, seed=5)      # T=64, L=2, 2 minibatches of 16 agents


def _oracle_rollout(tr, p0):
    from jaxrl_b200 import ops
    T, B, A = tr.T, tr.B, tr.cfg.action_dim
    gum = torch.stack([ops.gumbel_noise(torch.empty(B, A, device=DEV), tr.sample_seed, t) for t in range(T)]).cpu()
    return O.rollout_decode(p0, tr.cfg, tr.env.obs.cpu(), tr.env.reward.cpu(), gum), gum


def _minibatch(tr, i):
    b, Bm, T = tr.b, tr.Bm, tr.T
    sl = slice(i * Bm, (i + 1) * Bm)
    c = lambda t: t[sl].cpu()
    return dict(obs=c(b.obs), action_mask=None, actions=c(b.actions), last_actions=c(b.last_actions),
                last_rewards=c(b.last_rewards), terminated=torch.zeros(Bm, T, dtype=torch.bool), log_prob=c(b.log_prob),
                advantages=c(b.advantages), targets=c(b.targets))


@pytest.fixture(scope="module")
def stepped():
    """A trainer after warm-up (which must leave the seeded state untouched) and ONE train_step, optimizer off: the
    parameters are still the initial ones and p.grad holds the last minibatch's gradient."""
    from jaxrl_b200.trainer import PPOTrainer
    tr = PPOTrainer(RUN, use_graphs=True, run_optimizer=False)
    init = tr.engine.p.flat.clone()
    tr.warmup()
    assert torch.equal(tr.engine.p.flat, init), "warm-up must restore the parameters"
    tr.train_step()
    torch.cuda.synchronize()
    return tr


def test_rollout_matches_oracle(stepped):
    tr = stepped
    p0 = tr.engine.p.oracle_dict()
    ref, gum = _oracle_rollout(tr, p0)
    T, B = tr.T, tr.B
    act = tr.r.actions_ext[1:].t().cpu()
    same = act == ref["actions"]
    lead = same.int().cumprod(dim=1).bool()              # steps before an agent's first disagreement
    full = lead[:, -1].float().mean().item()
    _log(f"composed rollout: {lead.float().mean().item():.4f} of steps before the first disagreement, "
         f"{full:.3f} of agents identical over all {T} steps")
    # every first disagreement must be a near-tie of the oracle's perturbed logits (bf16 logits differ by <= 1e-2 rel)
    for b_ in range(B):
        if not lead[b_, -1]:
            t0 = int((~lead[b_]).nonzero()[0])
            z = ref["logits"][b_, t0] + gum[t0, b_]
            top = torch.topk(z, 2).values
            assert (top[0] - top[1]).item() < 5e-2, f"agent {b_} step {t0}: actions differ with margin {(top[0] - top[1]).item()}"
            assert int(act[b_, t0]) in torch.topk(z, 2).indices.tolist()
    assert full >= 0.5
    dv = (tr.r.values[:T].t().cpu() - ref["values"][:, :T]).abs()[lead].max().item()
    dl = (tr.r.log_prob.t().cpu() - ref["log_prob"]).abs()[lead].max().item()
    _log(f"composed rollout: max|dvalue|={dv:.3e} max|dlogp|={dl:.3e} (steps before divergence)")
    assert dv <= 2.5e-2 and dl <= 1e-2
    # bootstrap value (train.py:142-148 quirk: initial timestep, empty carry) - independent of the sampled actions
    db = (tr.r.values[T].cpu() - ref["values"][:, T]).abs().max().item()
    assert db <= 2e-2, f"bootstrap value error {db}"
    # first step: no history, no divergence possible
    assert (tr.r.values[0].cpu() - ref["values"][:, 0]).abs().max().item() <= 2e-2


def test_layout_and_gae_match_oracle(stepped):
    """Rollout.store semantics (rollout.py:102-126) + create_minibatches permutation (:169-181) + GAE (:128-167)."""
    tr = stepped
    b, r, env, T, B = tr.b, tr.r, tr.env, tr.T, tr.B
    perm = b.perm.cpu().long()
    assert sorted(perm.tolist()) == list(range(B))
    p0 = tr.engine.p.oracle_dict()
    ref, _ = _oracle_rollout(tr, p0)
    same = (r.actions_ext[1:].t().cpu() == ref["actions"]).all(dim=1)
    # env-driven fields: exact for every agent (the oracle's RolloutState rows, permuted)
    assert torch.equal(b.rewards.cpu(), ref["rewards"][perm])                 # rewards[:, t] = next_timestep.reward
    assert torch.equal(b.last_rewards.cpu(), ref["last_rewards"][perm])       # last_rewards[:, t] = timestep.reward
    assert torch.equal(b.obs.cpu(), env.obs[:T].cpu().transpose(0, 1)[perm])
    nt = torch.zeros(B, T, dtype=torch.uint8)
    nt[:, -1] = 1
    assert torch.equal(b.next_terminated.cpu(), nt)
    # action-driven fields: exact for the agents whose trajectories match the oracle's
    rows = same[perm]
    assert rows.any()
    assert torch.equal(b.actions.cpu()[rows], ref["actions"][perm][rows])
    assert torch.equal(b.last_actions.cpu()[rows], ref["last_actions"][perm][rows])
    # and for every agent against the kernel's own time-major rollout rows
    assert torch.equal(b.actions.cpu(), r.actions_ext[1:].t().cpu()[perm])
    assert torch.equal(b.last_actions.cpu(), r.actions_ext[:T].t().cpu()[perm])
    assert torch.equal(b.log_prob.cpu(), r.log_prob.t().cpu()[perm])
    assert torch.equal(b.values.cpu(), r.values.t().cpu()[perm])
    assert (b.last_actions[:, 0] == 0).all()
    # GAE on exactly these arrays
    hyp = tr.hyp
    adv_ref, tgt_ref = O.calculate_advantage(b.rewards.cpu().numpy(), b.values.cpu().numpy(),
                                             b.next_terminated.cpu().numpy().astype(bool), hyp.discount, hyp.gae_lambda,
                                             hyp.normalize_advantage)
    np.testing.assert_array_equal(b.targets.cpu().numpy(), tgt_ref)
    np.testing.assert_allclose(b.advantages.cpu().numpy(), adv_ref, rtol=1e-5, atol=1e-5)


def test_minibatch_gradient_matches_oracle(stepped):
    """Optimizer off: p.grad is the gradient of the LAST minibatch at the initial parameters."""
    tr = stepped
    m = tr.hyp.minibatch_count
    p = {k: v.clone().requires_grad_(True) for k, v in tr.engine.p.oracle_dict().items()}
    total, logs = O.ppo_loss(p, tr.cfg, tr.hyp, _minibatch(tr, m - 1), progress=0.0)
    total.backward()
    grads = tr.engine.p.grad_dict()
    for name, t in p.items():
        rel = (grads[name] - t.grad).norm().item() / (t.grad.norm().item() + 1e-8)
        _log(f"composed minibatch grad {name}: rel_l2={rel:.4f}")
        assert rel < 1.2e-2, f"{name}: rel l2 {rel}"
    got = tr.tws.losses.cpu()
    assert abs(got[0].item() - logs["value_loss"].item()) <= 1e-2 * abs(logs["value_loss"].item()) + 1e-3
    assert abs(got[1].item() - logs["actor_loss"].item()) <= 1e-2 * abs(logs["actor_loss"].item()) + 2e-3
    assert abs(got[2].item() - logs["entropy_loss"].item()) <= 1e-2 * abs(logs["entropy_loss"].item()) + 1e-3


@pytest.mark.parametrize("opt_type", ["muon", "adamw"])
def test_full_step_parameter_update_matches_oracle(opt_type):
    """The whole step with the optimizer on: the oracle repeats the m sequential (ppo_loss -> autograd -> optimizer)
    steps on the trainer's own batch-major arrays; compared: averaged minibatch losses and the parameter update.
    Both optimizers normalise the gradient (Newton-Schulz orthogonalisation / sign-like first Adam steps), which maps a
    0.5 % bf16 gradient error onto a few per cent of the update (measured: rel l2 4.2 % muon / 3.8 % adamw, cosine
    0.9993, norm ratio 1.000) - hence direction + norm + 8 % rel l2, not 1e-2 elementwise."""
    from jaxrl_b200.trainer import PPOTrainer
    run = replace(RUN, optimizer=replace(RUN.optimizer, type=opt_type))
    tr = PPOTrainer(run, use_graphs=True)
    tr.warmup()
    p0 = tr.engine.p.oracle_dict()
    assert int(tr.o.step.item()) == 0
    tr.train_step()
    torch.cuda.synchronize()
    m = tr.hyp.minibatch_count
    assert int(tr.o.step.item()) == m
    params = {k: v.clone() for k, v in p0.items()}
    state = {"mu": {k: torch.zeros_like(v) for k, v in params.items()}, "nu": {k: torch.zeros_like(v) for k, v in params.items()}}
    total_steps = run.optimizer_total_steps
    acc = dict(value_loss=0.0, actor_loss=0.0, entropy_loss=0.0)
    for i in range(m):
        pg = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        total, logs = O.ppo_loss(pg, tr.cfg, tr.hyp, _minibatch(tr, i), progress=0.0)
        total.backward()
        for k in acc:
            acc[k] += logs[k].item() / m
        params, _ = O.optimizer_step(params, {k: v.grad for k, v in pg.items()}, state, i, run.optimizer, total_steps,
                                     is_muon_param)
    got_logs = tr.logs()
    for k, v in acc.items():
        assert abs(got_logs[k] - v) <= 2e-2 * abs(v) + 2e-3, (k, got_logs[k], v)
    got = tr.engine.p.oracle_dict()
    d_ref = torch.cat([(params[k] - p0[k]).reshape(-1) for k in p0])
    d_got = torch.cat([(got[k] - p0[k]).reshape(-1) for k in p0])
    rel = ((d_got - d_ref).norm() / d_ref.norm()).item()
    cos = ((d_got * d_ref).sum() / (d_got.norm() * d_ref.norm())).item()
    nrm = (d_got.norm() / d_ref.norm()).item()
    _log(f"composed full step ({opt_type}): update rel_l2={rel:.4f} cos={cos:.5f} norm ratio={nrm:.4f}")
    assert cos > 0.998 and abs(nrm - 1) < 1e-2 and rel < 8e-2
