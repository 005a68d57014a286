#!/usr/bin/env python
"""Warm timings of the training GEMM shapes, with the epilogue partially disabled (mpx_debug_set(7, mode)) to attribute
time: mode 0 = normal, 1 = everything but the global stores, 2 = TMEM loads only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402
from jaxrl_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
M, H, F = 131072, 128, 768
bf = lambda *s: torch.randn(*s, device=dev, generator=g).to(torch.bfloat16)
x, hbuf = bf(M, H), bf(M, F)
wt_up = bf(F, H)            # (N, K)
wt_dn = bf(H, F)
wu, wg = torch.randn(H, F, device=dev, generator=g), torch.randn(H, F, device=dev, generator=g)
wt_ug = ops.pack_glu_weight(wu.cpu(), wg.cpu()).to(dev)
y768, y128 = torch.empty(M, F, dtype=torch.bfloat16, device=dev), torch.empty(M, H, dtype=torch.bfloat16, device=dev)
h, u, gg = (torch.empty(M, F, dtype=torch.bfloat16, device=dev) for _ in range(3))
bias = bf(F)
cases = {
    "dgrad K=128 -> N=768, plain store (201 MB out)": (lambda: ops.linear(x, wt_up, out=y768), 33.5 + 201.3),
    "vmlp  K=128 -> N=768, bias+gelu+y_pre (402 MB out)": (lambda: ops.linear(x, wt_up, bias=bias, act=1, out=y768, y_pre=h), 33.5 + 402.6),
    "down  K=768 -> N=128, +residual (201 MB in)": (lambda: ops.linear(hbuf, wt_dn, residual=x, out=y128), 201.3 + 33.5 * 2),
    "GLU up K=128 -> h,u,g (604 MB out)": (lambda: ops.glu_up(x, wt_ug, F, save=True, h=h, u=u, g=gg), 33.5 + 604),
}
for name, (fn, mb) in cases.items():
    line = f"{name:58s}"
    for mode in (0, 1, 2):
        lib.mpx_debug_set(7, mode)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        line += f"  mode{mode}: {us:7.1f} us" + (f" ({mb / us * 1e3 / 1e3:5.2f} TB/s)" if mode == 0 else "")
    lib.mpx_debug_set(7, 0)
    print(line)
