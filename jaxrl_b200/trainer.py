"""PPO update sequencing: rollout -> GAE -> minibatch forward/backward -> optimizer, mirroring the reference's
`train()` (mapox_trainer/train.py:225-309), `evaluate()` (:42-159) and `Rollout` (rollout.py) call for call, on top
of the C-ABI kernels.  Environments are replaced by synthetic observations of the named env shapes (north star);
the environment echoes the sampled action as next_timestep.last_action.

Data-parallel layout (new; the reference is single-device, train.py:312-315 is dead code): each rank owns
B/world agents end to end - KV caches, rollout buffers, GAE rows and minibatches - with weights replicated; per
minibatch one gradient all-reduce over NCCL and, once per rollout, an all-reduce of the advantage moments.
"""

from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import _lib, dist as mdist, ops
from .config import RunConfig
from .engine import PolicyEngine

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


class SyntheticEnv:
    """Pre-drawn observation / reward streams with the shapes of the named environment (T+1 steps, B agents).
    `host=True` keeps them in pinned host memory (the end-to-end path copies one step per decode step)."""

    def __init__(self, cfg, B: int, T: int, device, seed: int = 42, host: bool = False):
        gen = torch.Generator().manual_seed(seed)
        vw, vh = cfg.obs_shape[0], cfg.obs_shape[1]
        comps = [torch.randint(0, n, (T + 1, B, vw, vh), generator=gen, dtype=torch.uint8) for n in cfg.obs_max_value]
        # grid_cnn: (…, vw, vh, K) components; grid_flattened: one class id per cell, (…, vw, vh)
        obs = comps[0].contiguous() if cfg.obs_type == "grid_flattened" else torch.stack(comps, dim=-1).contiguous()
        reward = ((torch.rand(T + 1, B, generator=gen) < 0.02).float() * 0.05).contiguous()
        self.T, self.B = T, B
        self.obs_host = obs.pin_memory() if host else None
        self.reward_host = reward.pin_memory() if host else None
        self.obs = obs.to(device)
        self.reward = reward.to(device)
        term = torch.zeros(T + 1, B, dtype=torch.uint8)
        term[T] = 1                                   # episode ends with the rollout (max_env_steps == T)
        self.terminated = term.to(device)


class PPOTrainer:
    def __init__(self, run: RunConfig, device="cuda", rank: int = 0, world: int = 1, use_graphs: bool = True,
                 e2e_host_inputs: bool = False, seed: Optional[int] = None, run_optimizer: bool = True,
                 shard_optimizer: bool = True):
        self.run, self.cfg, self.hyp, self.opt = run, run.model, run.ppo, run.optimizer
        self.dev = torch.device(device)
        self.rank, self.world = rank, world
        self.use_graphs = use_graphs
        self.e2e = e2e_host_inputs
        self.run_optimizer = run_optimizer
        self.shard_optimizer = shard_optimizer            # world > 1: Muon sharded over the ranks (False: every rank repeats it)
        if run.num_agents % world != 0:
            raise ValueError("num_agents must divide evenly across ranks")
        self.B = run.num_agents // world
        self.T = self.cfg.max_seq_length
        m = self.hyp.minibatch_count
        if self.B % m != 0:
            raise ValueError(f"agents per rank ({self.B}) must be a multiple of minibatch_count ({m})")
        if self.hyp.epoch_count != 1:
            raise _lib.MpxError("epoch_count > 1 is not implemented on the B200 path yet")
        self.Bm = self.B // m
        seed = run.seed if seed is None else seed
        self.seed = seed
        self.sample_seed = seed + 104729 * rank
        # same seed on every rank: replicated weights; world > 1 lays the Muon matrices out in per-rank shards
        self.engine = PolicyEngine(self.cfg, self.dev, seed=seed, world=world if shard_optimizer else 1)
        mdist.broadcast_params(self.engine.p.flat)
        self.engine.stage_weights()
        self.env = SyntheticEnv(self.cfg, self.B, self.T, self.dev, seed=seed + 1000 * (rank + 1), host=self.e2e)
        self._alloc()
        self.rollout_graph = None
        self.update_graph = None

    # ------------------------------------------------------------------------------------------ buffers
    def _alloc(self):
        dev, B, T, cfg = self.dev, self.B, self.T, self.cfg
        A = cfg.action_dim
        self.dws = self.engine.alloc_decode(B)
        self.tws = self.engine.alloc_train(self.Bm, T)
        f = lambda *s: torch.zeros(*s, dtype=F32, device=dev)
        r = SimpleNamespace()
        # time-major rollout buffers (RolloutState, rollout.py:11-30, written one contiguous row per step)
        r.actions_ext = torch.zeros(T + 1, B, dtype=I32, device=dev)       # row 0 = last_action of the reset step
        r.log_prob = f(T, B)
        r.values = f(T + 1, B)
        r.gumbel = f(B, A)
        r.pos = torch.arange(T, dtype=I32, device=dev)[:, None].repeat(1, B).contiguous()   # ts.time per step
        # The value of step t is not an input of step t+1: the rollout only keeps the trunk output of every step and
        # evaluates the value path ONCE for all (T+1)*B rows afterwards (same arithmetic, off the per-step chain).
        self.defer_value = self.engine.fused_value
        r.hidden = torch.empty(T + 1, B, cfg.hidden_features, dtype=torch.bfloat16, device=dev) if self.defer_value else None
        # end-to-end mode: double-buffered device staging of the per-step host inputs
        r.obs_stage = [torch.empty_like(self.env.obs[0]) for _ in range(2)] if self.e2e else None
        r.rew_stage = [torch.empty_like(self.env.reward[0]) for _ in range(2)] if self.e2e else None
        self.copy_stream = torch.cuda.Stream(device=dev) if self.e2e else None
        self.r = r
        b = SimpleNamespace()
        # batch-major (permuted) training view
        b.obs = torch.empty(B, T, *self.env.obs.shape[2:], dtype=torch.uint8, device=dev)
        b.actions = torch.empty(B, T, dtype=I32, device=dev)
        b.last_actions = torch.empty(B, T, dtype=I32, device=dev)
        b.log_prob, b.rewards, b.last_rewards = f(B, T), f(B, T), f(B, T)
        b.values = f(B, T + 1)
        b.next_terminated = torch.empty(B, T, dtype=torch.uint8, device=dev)
        b.advantages, b.targets = f(B, T), f(B, T)
        b.stats = torch.zeros(2, dtype=torch.float64, device=dev)
        b.gae_ws = torch.empty(_lib.lib.mpx_gae_workspace_bytes(B, T), dtype=torch.uint8, device=dev)
        b.perm = torch.arange(B, dtype=I32, device=dev)
        self.b = b
        n = self.engine.p.numel
        o = SimpleNamespace(mu=f(n), nu=f(n), scratch=f(n), step=torch.zeros(1, dtype=I32, device=dev),
                            norm=f(1), ws=torch.empty(_lib.lib.mpx_adamw_workspace_bytes(n), dtype=torch.uint8, device=dev))
        if self.opt.type == "muon":
            # world > 1: sharded Muon - this rank runs Newton-Schulz on the matrices it owns only (optimizer.py:25-37,51-69)
            self.sharded = self.world > 1 and self.shard_optimizer and self.engine.p.muon_shard > 0
            tab, nmat, xt, at, mr, mc = self.engine.p.muon_table(self.rank if self.sharded else None)
            o.muon = SimpleNamespace(table=tab, nmat=nmat, x_total=xt, a_total=at, max_r=mr, max_c=mc,
                                     ns_ws=f(2 * xt + 2 * at + max(nmat, 1)),
                                     ws=torch.empty(_lib.lib.mpx_muon_workspace_bytes(n, max(nmat, 1)), dtype=torch.uint8,
                                                    device=dev))
        elif self.opt.type != "adamw":
            raise ValueError(f"Unknown optimizer type: {self.opt.type}")
        else:
            self.sharded = False
        self.o = o
        self.noise_off = torch.zeros(1, dtype=torch.int64, device=dev)      # Gumbel stream offset, += T per rollout
        self.ent_coef = torch.full((1,), self.hyp.entropy_coef, dtype=F32, device=dev)
        self.log_acc = f(4)
        self.update_idx = 0
        self.perm_gen = torch.Generator(device=dev).manual_seed(self.seed + 17 + self.rank)

    # ------------------------------------------------------------------------------------------ rollout (evaluate)
    def _rollout_chain(self, ws, rows: slice, seed: int):
        """The T decode steps (+ bootstrap value) of the agents in `rows` (evaluate(), train.py:42-159).
        End-to-end mode: every step's observations / rewards come from pinned host memory; the copy of step t+1 runs
        on a copy stream into the other half of a double buffer while step t computes."""
        eng, r, env, T = self.engine, self.r, self.env, self.T
        e2e = self.e2e
        if e2e:
            main = torch.cuda.current_stream()
            cs = self.copy_stream
            ready = [None, None]          # copy of the step using buffer b has landed (recorded on the copy stream)
            done = [None, None]           # the step that read buffer b has been enqueued completely (main stream)
            start = torch.cuda.Event()
            start.record(main)

            def fetch(t_src: int, b: int):
                with torch.cuda.stream(cs):
                    cs.wait_event(done[b] if done[b] is not None else start)
                    r.obs_stage[b][rows].copy_(env.obs_host[t_src][rows], non_blocking=True)
                    r.rew_stage[b][rows].copy_(env.reward_host[t_src][rows], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    ready[b] = ev
            fetch(0, 0)
        for t in range(T + 1):
            # step T is the bootstrap (train.py:142-148): value of the INITIAL timestep with the INITIAL (empty) carry.
            # Position 0 reads and writes ring slot 0 only, so re-running step 0 on the used caches is equivalent.
            src = t if t < T else 0
            if e2e:
                b = t & 1
                if t < T:
                    fetch(t + 1 if t + 1 < T else 0, b ^ 1)
                main.wait_event(ready[b])
                obs_t, rew_t = r.obs_stage[b][rows], r.rew_stage[b][rows]
            else:
                obs_t, rew_t = env.obs[src][rows], env.reward[src][rows]
            hid = r.hidden[t][rows] if self.defer_value else None
            val = None if self.defer_value else r.values[t][rows]
            if t < T:
                eng.decode_step(ws, obs_t, rew_t, r.actions_ext[t][rows], r.pos[t][rows], value_out=val, hidden_out=hid,
                                sample=dict(action=r.actions_ext[t + 1][rows], log_prob=r.log_prob[t][rows], seed=seed,
                                            stream_id=t, stream_offset=self.noise_off, want_logits=False))
            else:
                eng.decode_step(ws, obs_t, rew_t, r.actions_ext[0][rows], r.pos[0][rows], value_out=val, hidden_out=hid)
            if e2e:
                ev = torch.cuda.Event()
                ev.record(main)
                done[b] = ev

    def _rollout_body(self):
        self.engine.stage_rollout_weights()                 # weights are fixed for the whole rollout
        # the sampling key is per rank: the Philox counter is the LOCAL row, so equal seeds would hand agent i of every
        # shard the same Gumbel noise (perfectly correlated sampling across shards instead of i.i.d.)
        self._rollout_chain(self.dws, slice(0, self.B), self.sample_seed)
        self._deferred_values()

    def _deferred_values(self):
        if self.defer_value:
            r = self.r
            self.engine.values_from_hidden(r.hidden.view(-1, r.hidden.shape[-1]), r.values.view(-1))

    def _postprocess_body(self):
        """Layout change (+ row permutation, rollout.py:169-181) and calculate_advantage (rollout.py:128-167)."""
        r, b, env, T, B = self.r, self.b, self.env, self.T, self.B
        perm = b.perm
        ops.transpose_tb(env.obs[:T], b.obs, perm, T, B)
        ops.transpose_tb(r.actions_ext[1:], b.actions, perm, T, B)
        ops.transpose_tb(r.actions_ext[:T], b.last_actions, perm, T, B)
        ops.transpose_tb(r.log_prob, b.log_prob, perm, T, B)
        ops.transpose_tb(env.reward[1:], b.rewards, perm, T, B)
        ops.transpose_tb(env.reward[:T], b.last_rewards, perm, T, B)
        ops.transpose_tb(r.values, b.values, perm, T + 1, B)
        ops.transpose_tb(env.terminated[1:], b.next_terminated, perm, T, B)
        ops.gae(b.rewards, b.values, b.next_terminated, self.hyp.discount, self.hyp.gae_lambda, adv=b.advantages,
                tgt=b.targets, stats=b.stats, ws=b.gae_ws)

    def _normalize(self):
        b = self.b
        if not self.hyp.normalize_advantage:
            return
        count = mdist.allreduce_adv_moments(b.stats, self.B * self.T)
        ops.adv_normalize(b.advantages.view(-1), b.stats, count)

    # ------------------------------------------------------------------------------------------ update (train.py:249-260)
    def _minibatch_body(self, i: int):
        eng, b, tws, Bm = self.engine, self.b, self.tws, self.Bm
        sl = slice(i * Bm, (i + 1) * Bm)
        batch = dict(actions=b.actions[sl].reshape(-1), log_prob=b.log_prob[sl].reshape(-1),
                     advantages=b.advantages[sl].reshape(-1), targets=b.targets[sl].reshape(-1),
                     last_rewards=b.last_rewards[sl].reshape(-1), last_actions=b.last_actions[sl].reshape(-1),
                     obs=b.obs[sl])
        eng.p.grad.zero_()
        eng.forward_train(tws, b.obs[sl], batch["last_rewards"], batch["last_actions"])
        o, opt = self.o, self.opt
        total = self.run.optimizer_total_steps        # train.py:360-365
        n_adam = eng.p.n_adam

        def muon(phases: int, gscale: float):   # optimizer.py:25-37 (muon's own eps/beta defaults; adam_b1/b2 from the config)
            m = o.muon
            ops.muon_step(eng.p.flat, eng.p.grad, o.mu, o.nu, o.scratch, n_adam, m.table, m.nmat, m.x_total,
                          m.a_total, m.max_r, m.max_c, m.ns_ws, o.step, opt.learning_rate, total, opt.beta1, opt.beta2,
                          1e-8, opt.weight_decay, opt.max_norm, grad_scale=gscale, norm_out=o.norm, ws=m.ws, phases=phases)

        # The Muon-labelled matrices (everything but the encoders / value head) have their gradients before the
        # observation-encoder backward starts: their all-reduce and Newton-Schulz iterations run under it on the side
        # stream; the AdamW leaves, the global-norm clip and the parameter write follow the rest of the backward.
        early = self.run_optimizer and opt.type == "muon" and 0 < n_adam < eng.p.numel and \
            (o.muon.nmat > 0 or self.sharded)
        shard = eng.p.muon_shard

        def sharded_phase1():
            # reduce-scatter the Muon partition (each rank receives the summed gradient of ITS matrices), Newton-Schulz on
            # the local shard, all-gather the updates: 1/world of the orthogonalisation work and of the gradient traffic
            gs = mdist.reduce_scatter_shard(eng.p.grad[n_adam:], shard)
            if o.muon.nmat > 0:
                muon(1, gs)
            mdist.all_gather_shards(o.scratch[n_adam:], shard)

        if early and self.sharded:
            hook = sharded_phase1
        elif early:
            hook = lambda: muon(1, mdist.allreduce_gradients(eng.p.grad[n_adam:]))
        else:
            hook = None
        eng.loss_and_backward(tws, batch, self.hyp, 0.0, entropy_coef_dev=self.ent_coef, matrices_ready=hook)
        self.log_acc[:3] += tws.losses[:3]
        if early:
            gscale = mdist.allreduce_gradients(eng.p.grad[:n_adam])
            eng._wait(eng.matrices_event)
            if self.sharded:
                ops.muon_apply_sharded(eng.p.flat, eng.p.grad, o.mu, o.nu, o.scratch, n_adam, o.step, opt.learning_rate,
                                       total, opt.beta1, opt.beta2, 1e-8, opt.weight_decay, opt.max_norm, grad_scale=gscale,
                                       norm_out=o.norm, ws=o.muon.ws)
            else:
                muon(2, gscale)
            eng.stage_weights(rollout=False)
            return
        gscale = mdist.allreduce_gradients(eng.p.grad)
        if self.run_optimizer:
            if self.sharded:
                raise _lib.MpxError("sharded Muon needs both an AdamW and a Muon partition")
            if opt.type == "muon":
                muon(3, gscale)
            else:                         # optimizer.py:12-23
                ops.adamw_step(eng.p.flat, eng.p.grad, o.mu, o.nu, o.scratch, o.step, opt.learning_rate, total, opt.beta1,
                               opt.beta2, opt.eps, opt.weight_decay, opt.max_norm, grad_scale=gscale,
                               norm_out=o.norm, ws=o.ws)
            eng.stage_weights(rollout=False)

    def _update_body(self):
        for i in range(self.hyp.minibatch_count):
            self._minibatch_body(i)

    # ------------------------------------------------------------------------------------------ driver
    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def warmup(self):
        """One eager pass (sets kernel attributes, fills caches), then CUDA-graph capture of the three phases.
        The pass is a throw-away: parameters, optimizer moments and the device step counter are restored afterwards,
        so training starts from the seeded initial state (LR schedule and bias correction at step 0) and the first
        real rollout draws Gumbel stream 0."""
        keep = self.snapshot_state()
        self._rollout_body()
        self._postprocess_body()
        self._normalize()
        self._update_body()
        torch.cuda.synchronize()
        if self.use_graphs:
            self.rollout_graph = self._capture(self._rollout_body)
            self.post_graph = self._capture(self._postprocess_body)
            # world > 1: the update graph contains the NCCL gradient all-reduces
            self.update_graph = self._capture(self._update_body)
            torch.cuda.synchronize()
        self.restore_state(keep)
        self.log_acc.zero_()
        torch.cuda.synchronize()

    def snapshot_state(self):
        """Copies of everything an update mutates: parameters, optimizer moments, the device step counter."""
        o, p = self.o, self.engine.p
        return [t.clone() for t in (p.flat, o.mu, o.nu, o.step)]

    def restore_state(self, keep):
        o, p = self.o, self.engine.p
        for dst, src in zip((p.flat, o.mu, o.nu, o.step), keep):
            dst.copy_(src)
        self.engine.stage_weights()

    def train_step(self):
        """One outer PPO update (train.py:_global_step): returns the device tensor of summed minibatch losses."""
        progress = self.update_idx / max(1, self.run.update_steps)
        self.ent_coef.fill_(self.hyp.entropy_coef + progress * (self.hyp.entropy_coef_end - self.hyp.entropy_coef))
        self.b.perm.copy_(torch.randperm(self.B, generator=self.perm_gen, device=self.dev).to(I32))
        self.log_acc.zero_()
        if self.rollout_graph is not None:
            self.rollout_graph.replay()
            self.post_graph.replay()
        else:
            self._rollout_body()
            self._postprocess_body()
        self._normalize()
        if self.update_graph is not None:
            self.update_graph.replay()
        else:
            self._update_body()
        self.noise_off += self.T
        self.update_idx += 1
        return self.log_acc

    def timed_phases(self) -> dict:
        """One update with CUDA events between the phases (ms): rollout / layout+GAE / minibatch updates."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self.log_acc.zero_()
        ev[0].record()
        if self.rollout_graph is not None:
            self.rollout_graph.replay()
        else:
            self._rollout_body()
        ev[1].record()
        if self.rollout_graph is not None:
            self.post_graph.replay()
        else:
            self._postprocess_body()
        self._normalize()
        ev[2].record()
        if self.update_graph is not None:
            self.update_graph.replay()
        else:
            self._update_body()
        ev[3].record()
        torch.cuda.synchronize()
        self.noise_off += self.T
        self.update_idx += 1
        return {"rollout_ms": ev[0].elapsed_time(ev[1]), "gae_layout_ms": ev[1].elapsed_time(ev[2]),
                "update_ms": ev[2].elapsed_time(ev[3])}

    def logs(self) -> dict:
        """Device -> host read of the averaged minibatch losses (the reference's logger.log sync, train.py:419)."""
        n = self.hyp.minibatch_count * self.hyp.epoch_count
        v = (self.log_acc / n).cpu()
        return dict(value_loss=v[0].item(), actor_loss=v[1].item(), entropy_loss=v[2].item())
