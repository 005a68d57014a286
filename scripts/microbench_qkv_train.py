#!/usr/bin/env python
"""Warm timing of the training-shape q/k/v projection + RoPE GEMM (M = 131072 tokens, H = 128 -> 192) and of a plain
GEMM of the same shape (the difference is the RoPE epilogue)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402

dev = "cuda"
M, H, Nq, Nkv, D, T = 131072, 128, 4, 1, 32, 512
x = torch.randn(M, H, device=dev).to(torch.bfloat16)
w = (torch.randn(192, H, device=dev) / 11.3).to(torch.bfloat16)
qkv = torch.empty(M, 192, dtype=torch.bfloat16, device=dev)
pos = torch.arange(T, dtype=torch.int32, device=dev)
ts = (torch.tensor(10000.0) ** (2.0 * torch.arange(0, D // 2, dtype=torch.float32) / D)).to(dev)
y = torch.empty(M, 192, dtype=torch.bfloat16, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


t1 = timeit(lambda: ops.qkv_rope(x, w, pos, ts, Nq, Nkv, D, q=qkv, k=qkv[:, 128:], v=qkv[:, 160:], q_row_stride=192,
                                 k_row_stride=192, v_row_stride=192, pos_len=T))
t2 = timeit(lambda: ops.linear(x, w, out=y))
print(f"qkv_rope (M={M}): {t1:.1f} us    plain linear 128 -> 192: {t2:.1f} us")
