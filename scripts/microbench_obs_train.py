"""Warm timings of the training-size observation encoder (M = 256 x 512 tokens, 11x11x2 view): the engine's unfused
chain (conv1 gather kernel + im2col + GEMMs + embed_combine / dgrad GEMMs + col2im + colsums + weight gradients) against
the fused kernels of obs_train.cu, through PolicyEngine._embed_fwd / _embed_bwd."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jaxrl_b200.config import ModelConfig
from jaxrl_b200.engine import PolicyEngine

dev = "cuda"
Bm, T = 256, 512
M = Bm * T
cfg = ModelConfig(num_layers=2, max_seq_length=T, action_dim=8)
gen = torch.Generator().manual_seed(0)
obs = torch.stack([torch.randint(0, n, (M, 11, 11), generator=gen) for n in cfg.obs_max_value], -1).to(torch.uint8).to(dev)
reward = torch.randn(M, generator=gen).to(dev)
la = torch.randint(0, cfg.action_dim, (M,), generator=gen).to(torch.int32).to(dev)
dx0 = (torch.randn(M, cfg.hidden_features, generator=gen) * 0.01).to(torch.bfloat16).to(dev)


def timeit(fn, n=10):
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


res = {}
for fused in (False, True):
    eng = PolicyEngine(cfg, dev, seed=1, fused_obs_train=fused)
    ws = eng.alloc_train(Bm, T)
    tag = "fused" if fused else "unfused"
    res[f"forward {tag}"] = timeit(lambda: eng._embed_fwd(ws, M, obs, reward, la, None, ws.x0, train=True))
    res[f"backward {tag}"] = timeit(lambda: eng._embed_bwd(ws, M, dx0.clone(), reward, la, None, obs))
    del ws, eng
    torch.cuda.empty_cache()
for k, v in res.items():
    print(f"{k:24s} {v:8.1f} us")
print(json.dumps(res))
