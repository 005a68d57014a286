#!/usr/bin/env python
"""Phase timeline (CTA 0, last row tile, %globaltimer) of the training-size fused GLU kernels - bring-up aid."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402
from jaxrl_b200._lib import lib  # noqa: E402

BF16 = torch.bfloat16
dev = torch.device("cuda:0")
M, H, F = 131072, 128, 768
gen = torch.Generator().manual_seed(0)
xn = torch.randn(M, H, generator=gen).to(BF16).to(dev)
dy = torch.randn(M, H, generator=gen).to(BF16).to(dev)
wu = torch.randn(H, F, generator=gen) / math.sqrt(H)
wg = torch.randn(H, F, generator=gen) / math.sqrt(H)
wd = torch.randn(F, H, generator=gen) / math.sqrt(F)
wt_ug64 = ops.pack_glu_weight64(wu, wg).to(dev)
wt_d = wd.t().contiguous().to(BF16).to(dev)
out = torch.empty(M, H, dtype=BF16, device=dev)
h = torch.empty(M, F, dtype=BF16, device=dev)
dug = torch.empty(M, 2 * F, dtype=BF16, device=dev)
dxn = torch.empty(M, H, dtype=BF16, device=dev)
stamps = torch.zeros(64, dtype=torch.int64, device=dev)


def timeline(name, fn, labels):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    stamps.zero_()
    lib.mpx_debug_set_stamps(stamps.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.mpx_debug_set_stamps(None)
    s = stamps.cpu().tolist()
    t0 = s[0]
    print(f"--- {name}")
    for i, v in enumerate(s):
        if v:
            print(f"   slot {i:2d}  +{(v - t0) / 1e3:8.2f} us  {labels.get(i, '')}")


lab = {0: "start", 1: "setup done", 2: "x tile landed", 3: "xn ready", 20: "D ready", 23: "output stored", 24: "exit sync"}
for c in range(8):
    lab[4 + 2 * c] = f"ug{c} ready"
    lab[5 + 2 * c] = f"h{c} written"
for c in range(8):
    lab[32 + 3 * c] = f"[mma thread] before MMA1({c + 1})"
    lab[33 + 3 * c] = f"[mma thread] MMA1({c + 1}) issued"
    lab[34 + 3 * c] = f"[mma thread] h{c} seen, MMA2({c}) issued"
timeline("glu_fused fwd M=131072 (last tile of CTA 0)", lambda: ops.glu_fused(xn, wt_ug64, wt_d, xn, out, F, h_out=h), lab)
lab = {0: "start", 1: "setup done", 30: "chunks done", 31: "dxn ready", 32: "dxn stored"}
for c in range(8):
    lab[2 + 3 * c] = f"acc{c} ready"
    lab[3 + 3 * c] = f"math{c} done"
    lab[4 + 3 * c] = f"tiles{c} written"
timeline("glu_bwd_fused M=131072 (last tile of CTA 0)", lambda: ops.glu_bwd_fused(xn, dy, wt_ug64, wt_d, F, dug=dug, dxn=dxn), lab)
