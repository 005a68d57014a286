#!/usr/bin/env python
"""Warm, device-timed micro-benchmarks of the rollout-step kernels (one CUDA graph of N back-to-back launches per
variant; ncu's per-kernel durations are cold-cache and mislead for these latency-bound launches).
usage: python scripts/microbench_decode.py [--agents 4096] [--pos 300]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402
from jaxrl_b200.config import NAMED  # noqa: E402
from jaxrl_b200.engine import PolicyEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=4096)
ap.add_argument("--pos", type=int, default=300)
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--only", default=None, help="substring: run matching variants 3x without graphs (for ncu) and exit")
args = ap.parse_args()
dev = torch.device("cuda:0")
run = NAMED["return_2_layers"]
cfg = run.model
eng = PolicyEngine(cfg, dev, seed=0)
B = args.agents
ws = eng.alloc_decode(B)
eng.reset_carry(ws)
gen = torch.Generator(device=dev).manual_seed(0)
vw, vh, K = cfg.obs_shape
obs = torch.stack([torch.randint(0, n, (B, vw, vh), generator=gen, device=dev) for n in cfg.obs_max_value], -1).to(torch.uint8)
reward = torch.randn(B, device=dev, generator=gen)
la = torch.randint(0, cfg.action_dim, (B,), device=dev, generator=gen).to(torch.int32)
pos = torch.full((B,), args.pos, dtype=torch.int32, device=dev)
for kc, vc in zip(ws.kcache, ws.vcache):
    kc.normal_(generator=gen)
    vc.normal_(generator=gen)
p, w = eng.p, eng.w
S = cfg.max_seq_length
value = torch.empty(B, dtype=torch.float32, device=dev)
act = torch.empty(B, dtype=torch.int32, device=dev)
lp = torch.empty(B, dtype=torch.float32, device=dev)
gum = torch.empty(B, cfg.action_dim, dtype=torch.float32, device=dev)
x = torch.randn(B, eng.H, device=dev, generator=gen).to(torch.bfloat16)
x2 = torch.empty_like(x)
pre = "layers.0."


def obs_unfused():
    fused, eng.fused_obs = eng.fused_obs, False
    try:
        eng._embed_fwd(ws, B, obs, reward, la, None, ws.xa, train=False)
    finally:
        eng.fused_obs = fused


def heads_unfused():
    eng._heads_fwd(x, None, ws.xf, ws.logits, ws.pv, None, ws.vlogits)
    ops.hlgauss_value(ws.vlogits, eng.centers, out=value)
    ops.gumbel_noise(gum, 1, 2, None)
    ops.policy_sample(ws.logits, gum, action=act, log_prob=lp)


variants = {
    "obs_embed (unfused: conv1 + im2col + 2 GEMM + combine)": obs_unfused,
    "obs_embed_fused": lambda: eng._embed_fwd(ws, B, obs, reward, la, None, ws.xa, train=False),
    "rmsnorm": lambda: ops.rmsnorm_fwd(x, p[pre + "history_norm.scale"], out=ws.xn),
    "qkv_rope + cache write": lambda: ops.qkv_rope(ws.xn, w[pre + "wt_qkv"], pos, eng.timescale, eng.Nq, eng.Nkv, eng.D, q=ws.q,
                                                   k=ws.kcache[0], v=ws.vcache[0], k_row_stride=S * eng.KD,
                                                   v_row_stride=S * eng.KD, cache_S=S),
    "attn_decode": lambda: ops.attn_decode(ws.q, ws.kcache[0], ws.vcache[0], pos, eng.Nq, eng.Nkv, eng.D, out=ws.ao),
    "out_proj + residual": lambda: ops.linear(ws.ao, w[pre + "wt_o"], residual=x, out=x2),
    "glu_fused (norm inside)": lambda: ops.glu_fused(x, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F,
                                                     norm_scale=p[pre + "ffn_norm.scale"]),
    "glu_fused (pre-normalised)": lambda: ops.glu_fused(ws.xn, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F),
    "glu_fused + out projection (the rollout's feed-forward half)": lambda: ops.glu_fused(
        ws.ao, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F, norm_scale=p[pre + "ffn_norm.scale"], wt_o=w[pre + "wt_o"]),
    "glu_fused + out projection, then policy_head_sample (2 launches)": lambda: (
        ops.glu_fused(ws.ao, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F, norm_scale=p[pre + "ffn_norm.scale"],
                      wt_o=w[pre + "wt_o"]),
        ops.policy_head_sample(x2, p["action_embedder.embedding_table"], act, lp, seed=1, stream_id=2,
                               norm_scale=p["output_norm.scale"])),
    "glu_fused_head (feed-forward half + policy head, 1 launch)": lambda: ops.glu_fused_head(
        ws.ao, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F, p[pre + "ffn_norm.scale"], w[pre + "wt_o"],
        p["output_norm.scale"], p["action_embedder.embedding_table"], act, lp, seed=1, stream_id=2),
    "heads (unfused: norm, logits, vmlp, vhead, value, gumbel, sample)": heads_unfused,
    "policy_head_sample (fused)": lambda: ops.policy_head_sample(x, p["action_embedder.embedding_table"], act, lp, seed=1,
                                                                 stream_id=2, norm_scale=p["output_norm.scale"]),
    "value_head_fused": lambda: ops.value_head_fused(x, w["wt_vm"], w["b_vm"], w["wt_vh_hilo"], p["value_head.dense.bias"],
                                                     eng.centers, value, vlogits=ws.vlogits, norm_scale=p["output_norm.scale"]),
    "decode_step (whole)": lambda: eng.decode_step(ws, obs, reward, la, pos, value_out=value,
                                                   sample=dict(action=act, log_prob=lp, seed=1, stream_id=2, want_logits=False)),
}
if args.only:
    for name, fn in variants.items():
        if args.only in name:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
    sys.exit(0)
out = {}
for name, fn in variants.items():
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(args.reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (4 * args.reps)
    out[name] = round(us, 2)
    print(f"{us:9.2f} us  {name}", flush=True)
print(json.dumps({"microbench_decode_us": out, "agents": B, "pos": args.pos}))
