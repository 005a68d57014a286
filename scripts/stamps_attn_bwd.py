#!/usr/bin/env python
"""Phase timeline (CTA (0,0), thread 0, %globaltimer) of the tcgen05 attention backward kernel - bring-up aid.
Slots: 0 start; per iteration it < 6 and 64-key half: 1+10it+5half S/dP ready, 2+.. math done, 3+10it previous MMAs done
(p_empty), 4+.. P/dS half written, 5+10it dQ(it-1) drained (labelled "it half1 dQ drained")."""
import os, sys
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
ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], B, T, Nq, Nkv, D, QKV, QKV, QKV, o=o, lse=lse)
fn = lambda: ops.attn_train_bwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], o, dout, lse, pos, ts, B, T, Nq, Nkv, D, QKV, QKV, QKV,
                                dqkv, dqkv[:, QD:], dqkv[:, QD + KD:], QKV, QKV, QKV, delta=delta, dq_acc=dq_acc, pos_len=T)
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    fn()
b.record()
torch.cuda.synchronize()
print(f"attn bwd (delta + main + dq): {a.elapsed_time(b) / 20 * 1e3:.1f} us warm")
stamps = torch.zeros(64, dtype=torch.int64, device=dev)
lib.mpx_debug_set_stamps(stamps.data_ptr())
fn()
torch.cuda.synchronize()
lib.mpx_debug_set_stamps(None)
s = stamps.cpu().tolist()
names = {1: "S/dP ready", 2: "math done", 3: "prev MMAs done", 4: "tiles written", 0: "dQ(it-1) drained"}
t0 = min(v for v in s if v)
prev = t0
for i, v in sorted(((i, v) for i, v in enumerate(s) if v), key=lambda t: t[1]):
    lab = "start" if i == 0 else f"it{(i - 1) // 10} half{((i - 1) % 10) // 5} {names[i % 5]}"
    print(f"   slot {i:2d}  +{(v - t0) / 1e3:8.2f} us  (+{(v - prev) / 1e3:5.2f})  {lab}")
    prev = v


