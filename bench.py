#!/usr/bin/env python
"""PPO env-steps/s of the mapox_trainer transformer-over-time policy hot path on B200 (BASELINE.json metric).

One "step" = one PPO update of config/return_2_layers.json's shape on synthetic observations: rollout of T decode
steps for B agents (KV-cache decode + sampling + value), GAE, and minibatch_count forward/backward passes with the
gradient all-reduce and the optimizer step - i.e. the reference's `train()` (mapox_trainer/train.py:225-309).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config return_2_layers] [--impl ours|reference]
                  [--scaling weak|strong]

N > 1 is launched by torchrun (one rank per GPU, NCCL); agents are sharded across ranks.  `--scaling weak` (default):
B per GPU is the config's 4096; `--scaling strong`: the config's 4096 agents in total, 4096/N per GPU.  The default
(weak) line of an N > 1 run also carries a `strong_scaling` object measured in the same process.  Rank 0 prints ONE
JSON line.

`--impl reference` times the CPU restatement of the reference (oracle/, torch-CPU, all host threads) on a bounded
sample of the same workload; it imports nothing from `jaxrl_b200` (the CUDA library is never loaded on that arm).
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
CPU_SAMPLE_AGENTS = 64          # bounded CPU sample (both `cpu_baseline` and `--impl reference`)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="return_2_layers")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-extras", action="store_true",
                    help="no other_configs / strong_scaling / in-graph roofline legs (quick A/B runs)")
    ap.add_argument("--cpu-agents", type=int, default=None)
    ap.add_argument("--ncu", action="store_true", help="profiling run: one eager warm pass + ONE eager update, no extras")
    ap.add_argument("--ncu-decode-step", action="store_true", help="profiling run: one decode step (t=300) between profiler start/stop")
    ap.add_argument("--ncu-minibatch", action="store_true",
                    help="profiling run: eager warm pass, then ONE training minibatch (fwd+bwd+optimizer) between NVTX-free markers")
    ap.add_argument("--pdl", type=int, default=None,
                    help="experiment: programmatic-dependent-launch family mask (mpx_debug_set(6, mask); library default 12)")
    return ap.parse_args()


def workload_config(name: str, cfg, hyp, B: int, world: int, optimizer: str, scaling: str, graphs: bool) -> dict:
    """`config` of the JSON line - identical for both arms (the CPU arm's bounded sample is described in its
    `cpu_baseline.sample`, not here)."""
    T = cfg.max_seq_length
    return {"workload": f"{name}: PPO update = rollout of T={T} decode steps for B={B} agents/GPU + GAE + "
                        f"{hyp.minibatch_count} minibatch fwd/bwd + optimizer, L={cfg.num_layers} "
                        f"H={cfg.hidden_features} GQA {cfg.num_heads}:{cfg.num_kv_heads}",
            "agents_per_gpu": B, "trajectory": T, "layers": cfg.num_layers,
            "parallelism": f"env-sharded dp{world}", "optimizer": optimizer, "scaling": scaling,
            "l2": "working set (KV caches + activations, >2 GB) larger than the 126 MB L2; no flush",
            "cuda_graphs": graphs}


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


# ------------------------------------------------------------------------------------------------ CPU arm
def reference_arm(args):
    """The reference's CPU implementation of the path - here its torch-CPU restatement (oracle/; the JAX CPU backend
    cannot be installed in this image) - on all host threads, on a bounded sample of the SAME workload as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    cfg, hyp, B = synth.WORKLOADS[args.config]
    agents = args.cpu_agents or CPU_SAMPLE_AGENTS
    cb, dt = synth.time_cpu_update(args.config, agents, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.config, cfg, hyp, B if args.scaling == "weak" else B // args.gpus, args.gpus,
                                      "muon", args.scaling, True),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ ours
def count_launches(trainer) -> int:
    """Kernel launches of OUR library in one PPO update (counted from the C-ABI calls of one eager pass)."""
    from jaxrl_b200 import _lib, ops
    per_call = dict(ops.LAUNCHES_PER_CALL)
    counts = {"n": 0}
    orig = _lib.check

    def counting_check(rc, what=""):
        counts["n"] += per_call.get(what, 1)
        return orig(rc, what)

    ops.check = counting_check
    state = trainer.snapshot_state()
    try:
        trainer._rollout_body(); trainer._postprocess_body(); trainer._normalize(); trainer._update_body()
    finally:
        ops.check = orig
        trainer.restore_state(state)
    return counts["n"]


def _decode_attention_bytes(eng, B: int, T: int) -> float:
    """per agent-step per layer (T+1)/2 * Nkv*D*2*2 bytes of KV read + the rows in and out: x, out and
    the new k,v ring rows (fused; the projection weights come from L2) or q and out (unfused)."""
    kv_bytes = (T + 1) / 2 * eng.Nkv * eng.D * 2 * 2
    fused = eng.fused_attn_decode
    io_bytes = (eng.H * 2 + eng.Nq * eng.D * 2 + 2 * eng.Nkv * eng.D * 2) if fused else 2 * eng.Nq * eng.D * 2
    return B * (kv_bytes + io_bytes)


def decode_attention_roofline(trainer, peaks, in_graph: bool):
    """Dominant HBM-bound kernel: rollout decode attention (with the q/k/v projection folded in front where the
    shape allows).

    `frac` (in-graph): the kernel's time INSIDE the rollout graph = (replay time of the rollout graph - replay time of
    the same graph captured without the attention launches) / (T*L launches); everything else (encoder, feed-forward
    halves, policy head, value pass, programmatic dependent launches) is identical in both graphs.
    `frac_isolated`: the T launches of one rollout timed back to back with CUDA events on the launching stream
    (alternating layers so the 268 MB caches never sit in L2) - the kernel behind itself, the round-1 figure."""
    import torch
    from jaxrl_b200 import ops
    eng, r = trainer.engine, trainer.r
    ws = trainer.dws
    T, B, L = trainer.T, ws.B, eng.L
    fused = eng.fused_attn_decode
    x = torch.randn(B, eng.H, device=ws.ao.device).to(torch.bfloat16)

    def launch(t, i):
        if fused:
            pre = f"layers.{i}."
            ops.attn_decode_fused(x, eng.p[pre + "history_norm.scale"], eng.w[pre + "wt_qkv"], r.pos[t][:B], eng.timescale,
                                  ws.kcache[i], ws.vcache[i], eng.Nq, eng.Nkv, eng.D, out=ws.ao)
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
    iso_s = e0.elapsed_time(e1) * 1e-3 / T
    per_launch = _decode_attention_bytes(eng, B, T)
    peak = peaks.get("hbm_gbs", 6650.0)
    name = "attn_decode_fused_kernel" if fused else "attn_decode_online_kernel"
    out = {"kernel": name, "bound": "hbm", "peak": peak, "unit": "GB/s",
           "algorithmic_bytes_per_launch": per_launch, "agents_per_launch": B,
           "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
           "frac_isolated": per_launch / iso_s / 1e9 / peak, "isolated_launch_us": iso_s * 1e6}
    avg_s = iso_s
    if in_graph and trainer.rollout_graph is not None:
        def replay_ms(g, n=3):
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        full_ms = replay_ms(trainer.rollout_graph)
        key = "attn_decode_fused" if fused else "attn_decode"
        orig = getattr(ops, key)
        setattr(ops, key, lambda *a, **k: k.get("out"))           # bench-only: the same graph minus the attention launches
        try:
            g0 = trainer._capture(trainer._rollout_body)
        finally:
            setattr(ops, key, orig)
        without_ms = replay_ms(g0)
        del g0
        n_launch = (T + 1) * L
        avg_s = max(full_ms - without_ms, 1e-6) * 1e-3 / n_launch
        # the bootstrap step (position 0) moves no history: T*L launches carry the bytes
        out.update({"rollout_graph_ms": full_ms, "rollout_graph_without_kernel_ms": without_ms,
                    "timing": "in-graph marginal: (rollout graph - same graph without this kernel) / launches"})
        per_launch_graph = per_launch * T / (T + 1)
        out["achieved"] = per_launch_graph / avg_s / 1e9
    else:
        out["timing"] = "isolated back-to-back launches"
        out["achieved"] = per_launch / avg_s / 1e9
    out["frac"] = out["achieved"] / peak
    out["avg_launch_us"] = avg_s * 1e6
    # DRAM traffic per launch: ncu (dram__bytes_read.sum + dram__bytes_write.sum) / algorithmic bytes for this kernel at
    # one position, from the newest --set full capture under profiles/; scaled to the average launch like `achieved`.
    out["traffic"] = None
    for tfile in ("r02_attn_decode_fused_traffic.json", "r01_attn_decode_fused_traffic.json") if fused else ("r01_attn_decode_traffic.json",):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", tfile)))
            out["traffic"] = tj["ratio"] * per_launch
            out["traffic_source"] = f"profiles/{tfile} (ncu capture, ratio {tj['ratio']:.3f} of algorithmic)"
            break
        except Exception:
            continue
    return out


def train_flops_per_token(cfg) -> float:
    """per token, forward + backward: GEMMs 6 * params per layer, causal attention 3.5 * 4*Nq*D*(T+1)/2
    per layer, heads 6 * (H*Vh + Vh*NL + H*A) (the observation encoder is not counted)."""
    H, F, T = cfg.hidden_features, cfg.ffn_size, cfg.max_seq_length
    QD, KD = cfg.num_heads * cfg.head_dim, cfg.num_kv_heads * cfg.head_dim
    layer = 6.0 * (H * QD + 2 * H * KD + QD * H + 3 * H * F) + 3.5 * (4 * cfg.num_heads * cfg.head_dim * (T + 1) / 2)
    Vh = cfg.value_hidden_dim or 0
    heads = 6.0 * (H * Vh + (Vh or H) * cfg.value_n_logits + H * cfg.action_dim)
    return cfg.num_layers * layer + heads


def train_roofline(cfg, B: int, update_ms: float, peaks) -> dict:
    T = cfg.max_seq_length
    flops = train_flops_per_token(cfg) * B * T
    peak = peaks.get("bf16_tflops_sustained", 1355.0)
    achieved = flops / (update_ms * 1e-3) / 1e12
    out = {"phase": "PPO update (all minibatch fwd/bwd + optimizer)", "bound": "tensor", "achieved": achieved,
           "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "flops_per_update": flops,
           "flops_per_token": train_flops_per_token(cfg), "update_ms": update_ms,
           "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1355"}
    try:      # per-kernel tensor-pipe % of the newest ncu minibatch capture (profiles/, cold-cache, serialised)
        tp = json.load(open(os.path.join(ROOT, "profiles", "train_tensor_pct.json")))
        out["tensor_pipe_pct"] = tp["kernels"]
        out["tensor_pipe_source"] = tp["source"]
    except Exception:
        pass
    return out


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from dataclasses import replace
    from jaxrl_b200.config import NAMED
    from jaxrl_b200.trainer import PPOTrainer

    if args.pdl is not None:
        from jaxrl_b200 import _lib
        _lib.lib.mpx_debug_set(6, args.pdl)
    run = NAMED[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    T = run.model.max_seq_length
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass

    def sized(r, scaling):
        # weak: the config's agents on every GPU; strong: the config's agents in total
        return replace(r, num_agents=r.num_agents * world) if scaling == "weak" else r

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

    run_w = sized(run, args.scaling)
    B = run_w.num_agents // world

    if args.ncu_decode_step:
        trainer = PPOTrainer(run_w, rank=rank, world=world, use_graphs=False)
        trainer.warmup()
        t, r, env, eng = 300, trainer.r, trainer.env, trainer.engine
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        dws = trainer.dws
        rows = slice(0, dws.B)
        eng.decode_step(dws, env.obs[t][rows], env.reward[t][rows], r.actions_ext[t][rows], r.pos[t][rows],
                        hidden_out=r.hidden[t][rows] if trainer.defer_value else None,
                        value_out=None if trainer.defer_value else r.values[t][rows],
                        sample=dict(action=r.actions_ext[t + 1][rows], log_prob=r.log_prob[t][rows], seed=trainer.sample_seed,
                                    stream_id=t, stream_offset=trainer.noise_off, want_logits=False))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_decode_step": True}), flush=True)
        return
    if args.ncu_minibatch:
        trainer = PPOTrainer(run_w, rank=rank, world=world, use_graphs=False)
        trainer.warmup()
        trainer.train_step()
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
    graphs = not args.no_graphs
    trainer = PPOTrainer(run_w, rank=rank, world=world, use_graphs=graphs)
    trainer.warmup()
    launches = count_launches(trainer)      # every rank: the pass contains the collectives of one update
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    wu = max(3, args.warmup)
    secs = timed(trainer, args.steps, wu, sync_logs=False)
    clocks = sampler.stop() if rank == 0 else None
    logs = trainer.logs()
    value = world * B * T * args.steps / secs
    phases = trainer.timed_phases()
    roof = decode_attention_roofline(trainer, peaks, in_graph=not args.skip_extras) if rank == 0 else None
    roof_train = train_roofline(run.model, B, phases["update_ms"], peaks)
    del trainer
    torch.cuda.empty_cache()

    # ---- end-to-end arm: host (pinned) observations copied every decode step, losses read back every update
    e2e = None
    if not args.skip_e2e:
        tr2 = PPOTrainer(run_w, rank=rank, world=world, use_graphs=graphs, e2e_host_inputs=True)
        tr2.warmup()
        secs2 = timed(tr2, args.steps, wu, sync_logs=True)
        obs_bytes = tr2.env.obs_host[0].numel() * tr2.env.obs_host[0].element_size()
        rew_bytes = tr2.env.reward_host[0].numel() * 4
        e2e = {"value": world * B * T * args.steps / secs2, "unit": UNIT,
               "h2d_bytes_per_step": (T + 1) * (obs_bytes + rew_bytes), "d2h_bytes_per_step": 16}
        del tr2
        torch.cuda.empty_cache()

    # ---- extras: strong scaling in the same run (N > 1), the other BASELINE.json configs (N = 1)
    strong = None
    other = None
    if not args.skip_extras:
        def quick(r, steps=3, warm=3):
            tr = PPOTrainer(r, rank=rank, world=world, use_graphs=graphs)
            tr.warmup()
            s = timed(tr, steps, warm, sync_logs=False)
            ph = tr.timed_phases()
            Bq = r.num_agents // world
            del tr
            torch.cuda.empty_cache()
            return {"value": world * Bq * r.model.max_seq_length * steps / s, "unit": UNIT,
                    "ms_per_step": s / steps * 1e3, "agents_per_gpu": Bq, "steps": steps, "warmup": warm,
                    "phases_ms": ph}
        if world > 1 and args.scaling == "weak":
            strong = {}
            for name in (args.config, "return_8_layers"):
                r = NAMED[name]
                if r.num_agents % (world * r.ppo.minibatch_count) == 0:
                    strong[name] = quick(r)
        if world == 1 and args.config == "return_2_layers":
            other = {name: quick(NAMED[name]) for name in ("return_8_layers", "multitask", "blog_32_layers")}

    if rank == 0:
        cb = None
        if world == 1 and not args.skip_cpu:
            from oracle import synth
            cb, _ = synth.time_cpu_update(args.config, args.cpu_agents or CPU_SAMPLE_AGENTS)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": wu, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(run.name, run.model, run.ppo, B, world, run.optimizer.type, args.scaling, graphs),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
            "roofline": roof, "roofline_train": roof_train, "cpu_baseline": cb, "losses": logs, "phases_ms": phases,
            "strong_scaling": strong, "other_configs": other,
            "published_reference": "3.8M+ steps/s on 1x RTX 5090 (README.md:10), other hardware",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
