import os, sys, torch
sys.path.insert(0, "/root/repo")
from jaxrl_b200 import ops
from jaxrl_b200.config import NAMED
from jaxrl_b200.engine import PolicyEngine
dev = torch.device("cuda:0")
eng = PolicyEngine(NAMED["return_2_layers"].model, dev, seed=0)
M = 513 * 4096
x = torch.randn(M, 128, device=dev).to(torch.bfloat16)
val = torch.empty(M, dtype=torch.float32, device=dev)
for _ in range(2):
    eng.values_from_hidden(x, val)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.values_from_hidden(x, val)
e1.record()
torch.cuda.synchronize()
print("batched value pass over", M, "rows:", e0.elapsed_time(e1) / 5, "ms")
