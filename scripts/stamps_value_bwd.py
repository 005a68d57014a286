#!/usr/bin/env python
"""Phase timeline (CTA 0, last row tile, %globaltimer) of the fused value-path backward kernel - bring-up aid."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402
from jaxrl_b200._lib import lib  # noqa: E402

BF16 = torch.bfloat16
dev = torch.device("cuda:0")
M, H, Vh, NL = 131072, 128, 768, 51
gen = torch.Generator().manual_seed(0)
xf = torch.randn(M, H, generator=gen).to(BF16).to(dev)
dvl = torch.zeros(M, 64, dtype=BF16)
dvl[:, :NL] = (torch.randn(M, NL, generator=gen) * 0.02).to(BF16)
dvl = dvl.to(dev)
wvm = torch.randn(H, Vh, generator=gen) / math.sqrt(H)
wt_vm = wvm.t().contiguous().to(BF16).to(dev)
b_vm = (torch.randn(Vh, generator=gen) * 0.1).to(BF16).to(dev)
w_vh_pad = torch.zeros(Vh, 64, dtype=BF16)
w_vh_pad[:, :NL] = (torch.randn(Vh, NL, generator=gen) / math.sqrt(Vh)).to(BF16)
w_vh_pad = w_vh_pad.to(dev)
pv = torch.empty(M, Vh, dtype=BF16, device=dev)
dz = torch.empty(M, Vh, dtype=BF16, device=dev)
dxf = torch.empty(M, H, dtype=BF16, device=dev)
db = torch.zeros(Vh, device=dev)
ws = torch.empty(lib.mpx_value_bwd_fused_workspace_bytes(Vh), dtype=torch.uint8, device=dev)
stamps = torch.zeros(64, dtype=torch.int64, device=dev)
fn = lambda: ops.value_bwd_fused(xf, dvl, wt_vm, b_vm, w_vh_pad, pv, dz, dxf, db, ws=ws)
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    fn()
b.record()
torch.cuda.synchronize()
print(f"value_bwd_fused M={M}: {a.elapsed_time(b) / 10 * 1e3:.1f} us warm")
stamps.zero_()
lib.mpx_debug_set_stamps(stamps.data_ptr())
fn()
torch.cuda.synchronize()
lib.mpx_debug_set_stamps(None)
s = stamps.cpu().tolist()
lab = {0: "start", 1: "setup done", 30: "chunks done", 31: "dxf ready", 32: "dxf stored"}
for c in range(8):
    lab[2 + 3 * c] = f"acc{c} ready"
    lab[3 + 3 * c] = f"math{c} done"
    lab[4 + 3 * c] = f"tiles{c} written"
t0 = min(v for v in s if v)
for i, v in sorted(((i, v) for i, v in enumerate(s) if v), key=lambda t: t[1]):
    print(f"   slot {i:2d}  +{(v - t0) / 1e3:8.2f} us  {lab.get(i, '')}")
