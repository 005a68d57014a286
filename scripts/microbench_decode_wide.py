#!/usr/bin/env python
"""Warm device timings of the rollout-step launches of the hidden-256 configuration (multitask): the fused attention
half and the per-op feed-forward half (one CUDA graph of N back-to-back launches per variant).
usage: python scripts/microbench_decode_wide.py [--agents 4096] [--pos 300] [--pdl MASK]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import _lib, ops  # noqa: E402
from jaxrl_b200.config import NAMED  # noqa: E402
from jaxrl_b200.engine import PolicyEngine  # noqa: E402
from dataclasses import replace  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=4096)
ap.add_argument("--pos", type=int, default=300)
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--pdl", type=int, default=None, help="mpx_debug_set(6, mask): programmatic dependent launch families")
args = ap.parse_args()
if args.pdl is not None:
    _lib.lib.mpx_debug_set(6, args.pdl)
dev = torch.device("cuda:0")
cfg = replace(NAMED["multitask"].model, num_layers=2)
eng = PolicyEngine(cfg, dev, seed=0)
B = args.agents
ws = eng.alloc_decode(B)
eng.reset_carry(ws)
gen = torch.Generator(device=dev).manual_seed(0)
vw, vh, K = cfg.obs_shape
obs = torch.stack([torch.randint(0, n, (B, vw, vh), generator=gen, device=dev) for n in cfg.obs_max_value], -1).to(torch.uint8)
reward = torch.randn(B, device=dev, generator=gen)
la = torch.randint(0, cfg.action_dim, (B,), device=dev, generator=gen).to(torch.int32)
task = torch.randint(0, cfg.task_count, (B,), device=dev, generator=gen).to(torch.int32)
pos = torch.full((B,), args.pos, dtype=torch.int32, device=dev)
for kc, vc in zip(ws.kcache, ws.vcache):
    kc.normal_(generator=gen)
    vc.normal_(generator=gen)
p, w = eng.p, eng.w
S = cfg.max_seq_length
value = torch.empty(B, dtype=torch.float32, device=dev)
act = torch.zeros(B, dtype=torch.int32, device=dev)
lp = torch.empty(B, dtype=torch.float32, device=dev)
x = torch.randn(B, eng.H, device=dev, generator=gen).to(torch.bfloat16)
x2 = torch.empty_like(x)
pre = "layers.0."


def attn_fused(i=0):
    ops.attn_decode_fused(x, p[pre + "history_norm.scale"], w[pre + "wt_qkv"], pos, eng.timescale, ws.kcache[i & 1],
                          ws.vcache[i & 1], eng.Nq, eng.Nkv, eng.D, out=ws.ao)


def attn_chain(i=0):
    ops.rmsnorm_fwd(x, p[pre + "history_norm.scale"], out=ws.xn)
    ops.qkv_rope(ws.xn, w[pre + "wt_qkv"], pos, eng.timescale, eng.Nq, eng.Nkv, eng.D, q=ws.q, k=ws.kcache[i & 1],
                 v=ws.vcache[i & 1], k_row_stride=S * eng.KD, v_row_stride=S * eng.KD, cache_S=S)
    ops.attn_decode(ws.q, ws.kcache[i & 1], ws.vcache[i & 1], pos, eng.Nq, eng.Nkv, eng.D, out=ws.ao)


def out_proj(i=0):
    ops.linear(ws.ao, w[pre + "wt_o"], residual=x, out=x2)


def norm(i=0):
    ops.rmsnorm_fwd(x2, p[pre + "ffn_norm.scale"], out=ws.xn)


def glu_up(i=0):
    ops.glu_up(ws.xn, w[pre + "wt_ug"], eng.F, h=ws.h)


def down(i=0):
    ops.linear(ws.h, w[pre + "wt_d"], residual=x2, out=x)


def ffn_half(i=0):
    out_proj(); norm(); glu_up(); down()


def layer(i=0):
    attn_fused(i); ffn_half()


def step(i=0):
    eng.decode_step(ws, obs, reward, la, pos, task_ids=task, value_out=value,
                    sample=dict(action=act, log_prob=lp, seed=1, stream_id=i))


variants = [("attn_decode_fused (H=256)", attn_fused), ("rmsnorm + qkv_rope + attn_decode", attn_chain),
            ("out_proj + residual", out_proj), ("rmsnorm", norm), ("glu_up", glu_up), ("down + residual", down),
            ("feed-forward half (4 launches)", ffn_half), ("layer (5 launches)", layer),
            ("decode_step (2 layers, encoder, heads)", step)]
res = {}
for name, fn in variants:
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            for i in range(args.reps):
                fn(i)
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) * 1e3 / args.reps, 2)
    print(f"{res[name]:9.2f} us  {name}")
print(json.dumps({"microbench_decode_wide_us": res, "agents": B, "pos": args.pos, "pdl": args.pdl}))
