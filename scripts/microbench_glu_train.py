"""Warm timings of the training-size GLU block (M = 256 x 512 tokens, H = 128, F = 768): fused vs unfused chains."""
import json, math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jaxrl_b200 import ops

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
wt_ug = ops.pack_glu_weight(wu, wg).to(dev)
wt_d = wd.t().contiguous().to(BF16).to(dev)
w_d = wd.to(BF16).to(dev)
w_ug = torch.cat([wu, wg], dim=1).to(BF16).to(dev)
out = torch.empty(M, H, dtype=BF16, device=dev)
h = torch.empty(M, F, dtype=BF16, device=dev)
u = torch.empty(M, F, dtype=BF16, device=dev)
g = torch.empty(M, F, dtype=BF16, device=dev)
dh = torch.empty(M, F, dtype=BF16, device=dev)
dug = torch.empty(M, 2 * F, dtype=BF16, device=dev)
dxn = torch.empty(M, H, dtype=BF16, device=dev)


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


def fwd_unfused():
    ops.glu_up(xn, wt_ug, F, save=True, h=h, u=u, g=g)
    ops.linear(h, wt_d, residual=xn, out=out)


def bwd_unfused():
    ops.linear(dy, w_d, out=dh)
    ops.glu_bwd(dh, u, g, out=dug)
    ops.linear(dug, w_ug, out=dxn)


res = {
    "fwd unfused (glu_up save + down)": timeit(fwd_unfused),
    "fwd fused (glu_fused + h_out)": timeit(lambda: ops.glu_fused(xn, wt_ug64, wt_d, xn, out, F, h_out=h)),
    "bwd unfused (dh GEMM + glu_bwd + dxn GEMM)": timeit(bwd_unfused),
    "bwd fused (glu_bwd_fused)": timeit(lambda: ops.glu_bwd_fused(xn, dy, wt_ug64, wt_d, F, dug=dug, dxn=dxn)),
}
for k, v in res.items():
    print(f"{v:9.2f} us  {k}")
print(json.dumps({"microbench_glu_train_us": res, "M": M, "H": H, "F": F}))
