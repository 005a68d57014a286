#!/usr/bin/env python
"""Warm device timing of the rollout attention kernel as a function of the position (kv_len = pos + 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, S, Nq, Nkv, D = 4096, 512, 4, 1, 32
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(B, Nq * D, device=dev, generator=g).to(torch.bfloat16)
caches = [(torch.randn(B, S, Nkv, D, device=dev, generator=g).to(torch.bfloat16),
           torch.randn(B, S, Nkv, D, device=dev, generator=g).to(torch.bfloat16)) for _ in range(2)]
out = torch.empty_like(q)
H = int(sys.argv[sys.argv.index("--hidden") + 1]) if "--hidden" in sys.argv else 128
x = torch.randn(B, H, device=dev, generator=g).to(torch.bfloat16)
nscale = torch.ones(H, device=dev)
wt_qkv = (torch.randn((Nq + 2 * Nkv) * D, H, device=dev, generator=g) / H ** 0.5).to(torch.bfloat16)
half = D // 2
timescale = (10000.0 ** (torch.arange(half, dtype=torch.float32) * 2.0 / D)).to(dev)
fused = "--fused" in sys.argv


def launch(i, pos):
    if fused:
        ops.attn_decode_fused(x, nscale, wt_qkv, pos, timescale, caches[i & 1][0], caches[i & 1][1], Nq, Nkv, D, out=out)
    else:
        ops.attn_decode(q, caches[i & 1][0], caches[i & 1][1], pos, Nq, Nkv, D, out=out)


print(("fused norm+qkv+rope+attention" if fused else "attention only") + ": pos    us/launch   GB/s (algorithmic)")
for pos_v in (0, 15, 31, 63, 127, 191, 255, 319, 383, 447, 511):
    pos = torch.full((B,), pos_v, dtype=torch.int32, device=dev)
    for i in range(4):
        launch(i, pos)
    torch.cuda.synchronize()
    n = 40
    gr = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            for i in range(n):
                launch(i, pos)   # alternate caches: no L2 reuse
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    nbytes = B * ((pos_v + 1) * Nkv * D * 2 * 2 + 2 * Nq * D * 2)
    print(f"{pos_v:4d}   {us:9.2f}   {nbytes / us / 1e3:8.1f}")
