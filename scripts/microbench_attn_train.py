#!/usr/bin/env python
"""Warm device timings (CUDA events, 20 launches back to back) of the training attention kernels at the bench shape
(B=256 rows x T=512, 4 query heads on 1 kv head, head_dim 32) for every implementation selectable with
mpx_debug_set(2, impl)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402
from jaxrl_b200._lib import lib  # noqa: E402

B, T, Nq, Nkv, D = 256, 512, 4, 1, 32
QD, KD = Nq * D, Nkv * D
QKV = QD + 2 * KD
dev = "cuda"
qkv = torch.randn(B * T, QKV, device=dev).to(torch.bfloat16)
dout = torch.randn(B * T, QD, device=dev).to(torch.bfloat16)
dqkv = torch.zeros(B * T, QKV, dtype=torch.bfloat16, device=dev)
pos = torch.arange(T, dtype=torch.int32, device=dev)
ts = (torch.tensor(10000.0) ** (2.0 * torch.arange(0, D // 2, dtype=torch.float32) / D)).to(dev)
o = torch.empty(B * T, QD, dtype=torch.bfloat16, device=dev)
lse = torch.empty(B * T, Nq, dtype=torch.float32, device=dev)
delta = torch.empty(B * T, Nq, dtype=torch.float32, device=dev)
dq_acc = torch.empty(B * T, QD, dtype=torch.float32, device=dev)
flops_fwd = 4 * Nq * D * (T + 1) / 2 * B * T          # causal QK^T + PV


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


for impl in (0, 2, 1):
    lib.mpx_debug_set(2, impl)
    fwd = timeit(lambda: ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], B, T, Nq, Nkv, D, QKV, QKV, QKV, o=o, lse=lse))
    bwd = timeit(lambda: ops.attn_train_bwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], o, dout, lse, pos, ts, B, T, Nq, Nkv, D, QKV, QKV,
                                            QKV, dqkv, dqkv[:, QD:], dqkv[:, QD + KD:], QKV, QKV, QKV, delta=delta, dq_acc=dq_acc,
                                            pos_len=T))
    print(f"impl {impl}: fwd {fwd:7.1f} us ({flops_fwd / fwd / 1e6:6.1f} TFLOP/s)   bwd (3 launches) {bwd:7.1f} us "
          f"({2.5 * flops_fwd / bwd / 1e6:6.1f} TFLOP/s)")
lib.mpx_debug_set(2, 0)
