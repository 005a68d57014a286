#!/usr/bin/env python
"""Phase timeline (CTA 0, its last tile, %globaltimer) of the fused training observation encoder kernels - bring-up aid."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200.config import ModelConfig  # noqa: E402
from jaxrl_b200.engine import PolicyEngine  # noqa: E402
from jaxrl_b200._lib import lib  # noqa: E402

dev = "cuda"
Bm, T = 256, 512
M = Bm * T
cfg = ModelConfig(num_layers=2, max_seq_length=T, action_dim=8)
gen = torch.Generator().manual_seed(0)
obs = torch.stack([torch.randint(0, n, (M, 11, 11), generator=gen) for n in cfg.obs_max_value], -1).to(torch.uint8).to(dev)
reward = torch.randn(M, generator=gen).to(dev)
la = torch.randint(0, cfg.action_dim, (M,), generator=gen).to(torch.int32).to(dev)
dx0 = (torch.randn(M, cfg.hidden_features, generator=gen) * 0.01).to(torch.bfloat16).to(dev)
eng = PolicyEngine(cfg, dev, seed=1)
ws = eng.alloc_train(Bm, T)
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


timeline("obs_train_fwd (last tile of CTA 0)", lambda: eng._embed_fwd(ws, M, obs, reward, la, None, ws.x0, train=True),
         {0: "tile start", 1: "codes written", 2: "conv1 gather done", 3: "conv2 accumulators ready", 4: "z2 / a2 written",
          5: "embedding addends loaded", 6: "conv3 accumulator ready", 7: "staged", 8: "x stored"})
if len(sys.argv) > 1 and sys.argv[1] == "bwd":
    timeline("obs_train_bwd (last tile of CTA 0)", lambda: eng._embed_bwd(ws, M, dx0.clone(), reward, la, None, obs),
             {0: "tile start", 1: "dx0 tile in smem", 2: "da2 ready", 3: "dz2 written", 10: "rows done"})
