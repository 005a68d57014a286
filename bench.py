#!/usr/bin/env python
"""PPO env-steps/s of the mapox_trainer transformer-over-time policy hot path on B200 (BASELINE.json metric).

One "step" = one PPO update of config/return_2_layers.json's shape on synthetic observations: rollout of T decode
steps for B agents (KV-cache decode + sampling + value), GAE, and minibatch_count forward/backward passes with the
gradient all-reduce and the optimizer step - i.e. the reference's `train()` (mapox_trainer/train.py:225-309).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config return_2_layers] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); agents are sharded across ranks (weak scaling: B per GPU
is fixed).  Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ppo_env_steps_per_sec"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="return_2_layers")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--cpu-agents", type=int, default=None)
    ap.add_argument("--ncu", action="store_true", help="profiling run: one eager warm pass + ONE eager update, no extras")
    ap.add_argument("--ncu-decode-step", action="store_true", help="profiling run: one decode step (t=300) between profiler start/stop")
    ap.add_argument("--ncu-minibatch", action="store_true",
                    help="profiling run: eager warm pass, then ONE training minibatch (fwd+bwd+optimizer) between NVTX-free markers")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(run, agents: int, steps: int = 1, warmup: int = 0):
    """The oracle (torch-CPU restatement of the reference path; JAX cannot be installed here) timed on the host
    cores for the same unit of work on a bounded sample: `agents` agents x T steps, same L/H/minibatch structure."""
    import torch
    from oracle import oracle as O
    from jaxrl_b200.params import param_specs, _init

    cfg, hyp = run.model, run.ppo
    T = cfg.max_seq_length
    torch.set_num_threads(os.cpu_count() or 1)
    gen = torch.Generator().manual_seed(0)
    p = {name: _init(shape, kind, gen) for name, shape, kind in param_specs(cfg)}
    vw, vh = cfg.obs_shape[0], cfg.obs_shape[1]
    comps = [torch.randint(0, n, (T + 1, agents, vw, vh), generator=gen) for n in cfg.obs_max_value]
    obs = (comps[0] if cfg.obs_type == "grid_flattened" else torch.stack(comps, -1)).to(torch.uint8)
    reward = (torch.rand(T + 1, agents, generator=gen) < 0.02).float() * 0.05
    u = torch.rand(T, agents, cfg.action_dim, generator=gen).clamp(1e-6, 1 - 1e-6)
    gumbel = -torch.log(-torch.log(u))
    m = max(1, min(hyp.minibatch_count, agents))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            ro = O.rollout_decode(p, cfg, obs, reward, gumbel)
        term = torch.zeros(agents, T, dtype=torch.bool); term[:, -1] = True
        adv, tgt = O.calculate_advantage(ro["rewards"].numpy(), ro["values"].numpy(), term.numpy(), hyp.discount,
                                         hyp.gae_lambda, hyp.normalize_advantage)
        adv, tgt = torch.from_numpy(adv), torch.from_numpy(tgt)
        pg = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        bm = agents // m
        for i in range(m):
            sl = slice(i * bm, (i + 1) * bm)
            batch = dict(obs=obs[:T].transpose(0, 1)[sl], action_mask=None, actions=ro["actions"][sl],
                         last_actions=ro["last_actions"][sl], last_rewards=ro["last_rewards"][sl],
                         terminated=torch.zeros(bm, T, dtype=torch.bool), log_prob=ro["log_prob"][sl],
                         advantages=adv[sl], targets=tgt[sl])
            total, _ = O.ppo_loss(pg, cfg, hyp, batch, 0.0)
            total.backward()
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times[warmup:])
    return dict(value=agents * T / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{agents} agents x {T} steps x {cfg.num_layers} layers, {m} minibatches, torch-CPU oracle "
                       f"(restatement of the reference; JAX CPU backend not installable here), {dt:.2f} s per update"), dt


def reference_arm(args, run):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.cpu_agents = args.cpu_agents or 16
    cb, dt = cpu_baseline(run, args.cpu_agents, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{run.name}: PPO update (rollout T={run.model.max_seq_length} + GAE + "
                                   f"{run.ppo.minibatch_count} minibatch fwd/bwd), L={run.model.num_layers}, "
                                   f"bounded CPU sample of {args.cpu_agents} agents"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ ours
def count_launches(trainer) -> int:
    """Kernel launches of OUR library in one PPO update (counted from the C-ABI calls of one eager pass)."""
    from jaxrl_b200 import _lib
    # muon: prepare + 3 x 5 Newton-Schulz GEMMs + finish (phase 1), AdamW leaves + clip/apply (phase 2)
    per_call = {"mpx_gae": 2, "mpx_hlgauss_loss": 2, "mpx_ppo_policy_loss": 2, "mpx_adamw_step": 2,
                "mpx_attn_train_bwd": 3, "mpx_muon_step:1": 17, "mpx_muon_step:2": 2, "mpx_muon_step:3": 19}
    counts = {"n": 0}
    orig = _lib.check

    def counting_check(rc, what=""):
        counts["n"] += per_call.get(what, 1)
        return orig(rc, what)

    from jaxrl_b200 import ops
    ops.check = counting_check
    try:
        trainer._rollout_body(); trainer._postprocess_body(); trainer._normalize(); trainer._update_body()
    finally:
        ops.check = orig
    return counts["n"]


def decode_attention_roofline(trainer, peaks):
    """Dominant HBM-bound kernel: rollout decode attention (with the q/k/v projection folded in front where the
    shape allows).  Times the T launches of one rollout (positions 0..T-1, alternating layers so the 268 MB caches
    never sit in L2) back to back with CUDA events on the launching stream; achieved = algorithmic bytes per launch /
    average launch duration."""
    import torch
    from jaxrl_b200 import ops
    eng, r = trainer.engine, trainer.r
    ws = trainer.dws if trainer.n_chains == 1 else trainer.dws_chain[0]   # the launch shape the rollout really uses
    T, B, L = trainer.T, ws.B, eng.L
    fused = eng.fused_attn_decode
    x = torch.randn(B, eng.H, device=ws.ao.device).to(torch.bfloat16)

    def launch(t, i):
        if fused:
            pre = f"layers.{i}."
            ops.attn_decode_fused(x, eng.p[pre + "history_norm.scale"], eng.w[pre + "wt_qkv"], r.pos[t][:B], eng.timescale,
                                  ws.kcache[i], ws.vcache[i], eng.Nq, eng.Nkv, eng.D, out=ws.ao,
                                  stable_history=eng.early_kv_stream)    # as the rollout launches it
        else:
            ops.attn_decode(ws.q, ws.kcache[i], ws.vcache[i], r.pos[t][:B], eng.Nq, eng.Nkv, eng.D, out=ws.ao)

    for t in range(0, T, 64):
        launch(t, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(T):
        launch(t, t % L)
    e1.record()
    torch.cuda.synchronize()
    avg_s = e0.elapsed_time(e1) * 1e-3 / T
    # per agent-step per layer (T+1)/2 * Nkv*D*2*2 bytes of KV read + the rows in and out: q and out
    # (unfused) or x, out and the new k,v ring rows (fused; the 48 KB of projection weights per CTA come from L2)
    kv_bytes = (T + 1) / 2 * eng.Nkv * eng.D * 2 * 2
    io_bytes = (eng.H * 2 + eng.Nq * eng.D * 2 + 2 * eng.Nkv * eng.D * 2) if fused else 2 * eng.Nq * eng.D * 2
    per_launch = B * (kv_bytes + io_bytes)
    achieved = per_launch / avg_s / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM traffic per launch: ncu measured (read + write) / algorithmic bytes for this kernel at one position
    # (profiles/r01_attn_decode*_traffic.json); scaled to the average launch of the rollout, like `achieved`.
    traffic = None
    name = "attn_decode_fused_kernel" if fused else "attn_decode_online_kernel"
    tfile = "r01_attn_decode_fused_traffic.json" if fused else "r01_attn_decode_traffic.json"
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", tfile)))
        traffic = tj["ratio"] * per_launch
    except Exception:
        pass
    return {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "avg_launch_us": avg_s * 1e6,
            "algorithmic_bytes_per_launch": per_launch, "agents_per_launch": B,
            "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"}


def main():
    args = parse()
    from jaxrl_b200.config import NAMED
    run = NAMED[args.config]
    if args.impl == "reference":
        reference_arm(args, run)
        return
    import torch
    import torch.distributed as dist
    from dataclasses import replace
    from jaxrl_b200.trainer import PPOTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    # weak scaling: B agents per GPU fixed (the config's 4096), total = world * B
    run_w = replace(run, num_agents=run.num_agents * world)
    B, T = run.num_agents, run.model.max_seq_length
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(trainer, steps, warmup, sync_logs):
        for _ in range(warmup):
            trainer.train_step()
            if sync_logs:
                trainer.logs()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            trainer.train_step()
            if sync_logs:
                trainer.logs()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) * 1e-3
        t = torch.tensor([dev if not sync_logs else max(dev, wall)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    if args.ncu_decode_step:
        trainer = PPOTrainer(run_w, rank=rank, world=world, use_graphs=False)
        trainer.warmup()
        from jaxrl_b200 import ops as _ops
        t, r, env, eng = 300, trainer.r, trainer.env, trainer.engine
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        dws = trainer.dws if trainer.n_chains == 1 else trainer.dws_chain[0]
        rows = slice(0, dws.B)
        eng.decode_step(dws, env.obs[t][rows], env.reward[t][rows], r.actions_ext[t][rows], r.pos[t][rows],
                        value_out=r.values[t][rows],
                        sample=dict(action=r.actions_ext[t + 1][rows], log_prob=r.log_prob[t][rows], seed=trainer.seed,
                                    stream_id=t, stream_offset=trainer.noise_off, want_logits=False))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_decode_step": True}), flush=True)
        return
    if args.ncu_minibatch:
        trainer = PPOTrainer(run_w, rank=rank, world=world, use_graphs=False)
        trainer.warmup()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trainer._minibatch_body(1)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_minibatch": True}), flush=True)
        return
    if args.ncu:
        trainer = PPOTrainer(run_w, rank=rank, world=world, use_graphs=False)
        trainer.warmup()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()          # `ncu --profile-from-start off`: exactly one PPO update (= one bench step)
        trainer.train_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_run": True, "losses": trainer.logs()}), flush=True)
        return
    # ---- device-resident arm (value)
    trainer = PPOTrainer(run_w, rank=rank, world=world, use_graphs=not args.no_graphs)
    trainer.warmup()
    launches = count_launches(trainer)      # every rank: the pass contains the collectives of one update
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    secs = timed(trainer, args.steps, max(3, args.warmup), sync_logs=False)
    clocks = sampler.stop() if rank == 0 else None
    logs = trainer.logs()
    value = world * B * T * args.steps / secs
    roof = decode_attention_roofline(trainer, peaks) if rank == 0 else None
    phases = trainer.timed_phases()
    del trainer
    torch.cuda.empty_cache()

    # ---- end-to-end arm: host (pinned) observations copied every decode step, losses read back every update
    e2e = None
    if not args.skip_e2e:
        tr2 = PPOTrainer(run_w, rank=rank, world=world, use_graphs=not args.no_graphs, e2e_host_inputs=True)
        tr2.warmup()
        secs2 = timed(tr2, args.steps, max(3, args.warmup), sync_logs=True)
        obs_bytes = tr2.env.obs_host[0].numel() * tr2.env.obs_host[0].element_size()
        rew_bytes = tr2.env.reward_host[0].numel() * 4
        e2e = {"value": world * B * T * args.steps / secs2, "unit": UNIT,
               "h2d_bytes_per_step": (T + 1) * (obs_bytes + rew_bytes), "d2h_bytes_per_step": 16}
        del tr2
        torch.cuda.empty_cache()

    if rank == 0:
        cb = None
        if world == 1 and not args.skip_cpu:
            cb, _ = cpu_baseline(run, args.cpu_agents or 64)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{run.name}: PPO update = rollout of T={T} decode steps for B={B} agents/GPU + GAE + "
                                   f"{run.ppo.minibatch_count} minibatch fwd/bwd + optimizer, L={run.model.num_layers} "
                                   f"H={run.model.hidden_features} GQA {run.model.num_heads}:{run.model.num_kv_heads}",
                       "agents_per_gpu": B, "trajectory": T, "layers": run.model.num_layers,
                       "parallelism": f"env-sharded dp{world}", "optimizer": run.optimizer.type,
                       "l2": "working set (KV caches + activations, >2 GB) larger than the 126 MB L2; no flush",
                       "cuda_graphs": not args.no_graphs},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
            "roofline": roof, "cpu_baseline": cb, "losses": logs, "phases_ms": phases,
            "published_reference": "3.8M+ steps/s on 1x RTX 5090 (README.md:10), other hardware",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
