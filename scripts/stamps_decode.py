#!/usr/bin/env python
"""Phase timeline (CTA 0, %globaltimer) of the fused rollout kernels - bring-up aid, see mpx_debug_set_stamps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200 import ops  # noqa: E402
from jaxrl_b200._lib import lib  # noqa: E402
from jaxrl_b200.config import NAMED  # noqa: E402
from jaxrl_b200.engine import PolicyEngine  # noqa: E402

dev = torch.device("cuda:0")
cfg = NAMED["return_2_layers"].model
eng = PolicyEngine(cfg, dev, seed=0)
B = 4096
ws = eng.alloc_decode(B)
eng.reset_carry(ws)
gen = torch.Generator(device=dev).manual_seed(0)
vw, vh, K = cfg.obs_shape
obs = torch.stack([torch.randint(0, n, (B, vw, vh), generator=gen, device=dev) for n in cfg.obs_max_value], -1).to(torch.uint8)
reward = torch.randn(B, device=dev, generator=gen)
la = torch.randint(0, cfg.action_dim, (B,), device=dev, generator=gen).to(torch.int32)
x = torch.randn(B, eng.H, device=dev, generator=gen).to(torch.bfloat16)
x2 = torch.empty_like(x)
p, w = eng.p, eng.w
pre = "layers.0."
stamps = torch.zeros(64, dtype=torch.int64, device=dev)


def timeline(name, fn, labels):
    for _ in range(5):
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
            print(f"   slot {i:2d}  +{(v - t0) / 1e3:7.2f} us  {labels.get(i, '')}")


timeline("obs_fused", lambda: eng._embed_fwd(ws, B, obs, reward, la, None, ws.xa, train=False),
         {0: "start", 1: "setup done (barriers, tmem alloc)", 2: "obs tile in smem", 3: "codes done, image landed", 4: "conv1 gather done",
          5: "conv2 MMAs done", 6: "conv2 epilogue done", 7: "conv3 MMAs done", 8: "staging synced", 9: "embedding sum stored"})
timeline("glu_fused (norm)", lambda: ops.glu_fused(x, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F,
                                                   norm_scale=p[pre + "ffn_norm.scale"]),
         {0: "start", 1: "setup done (barriers, tmem, cluster sync)", 2: "x tile landed", 3: "norm done", 4: "ug0 ready", 5: "h0 written",
          6: "ug1 ready", 7: "h1 written", 8: "ug2 ready", 9: "h2 written", 20: "D ready", 21: "partial staged", 22: "all partials in",
          23: "output stored", 24: "exit sync"})
value = torch.empty(B, dtype=torch.float32, device=dev)
timeline("value_head_fused (norm)", lambda: ops.value_head_fused(x, w["wt_vm"], w["b_vm"], w["wt_vh_hilo"], p["value_head.dense.bias"],
                                                               eng.centers, value, vlogits=ws.vlogits,
                                                               norm_scale=p["output_norm.scale"]),
         {0: "start", 1: "setup done", 2: "x tile landed", 3: "norm done", 4: "z0 ready", 5: "h0 written", 6: "z1 ready", 7: "h1 written",
          8: "z2 ready", 9: "h2 written", 20: "D ready", 21: "partial staged", 22: "all partials in", 23: "row group summed",
          24: "values stored", 25: "exit"})

# training attention forward (tcgen05): CTA 0 = (last query tile, head 0, batch row 0): 8 key tiles
Bm, T = 256, 512
qkv = torch.randn(Bm * T, eng.QKV, device=dev, generator=gen).to(torch.bfloat16)
o = torch.empty(Bm * T, eng.QD, dtype=torch.bfloat16, device=dev)
lse = torch.empty(Bm * T, eng.Nq, dtype=torch.float32, device=dev)
QD, KD = eng.QD, eng.KD
lab = {0: "start", 1: "setup done", 30: "O / lse stored"}
for kt in range(8):
    lab[2 + 3 * kt] = f"S{kt} ready"
    lab[3 + 3 * kt] = f"MMA2({kt - 1}) done (P free)"
    lab[4 + 3 * kt] = f"P{kt} written + sync"
timeline("attn_fwd_tc (query tile 3, 8 key tiles)",
         lambda: ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], Bm, T, eng.Nq, eng.Nkv, eng.D, eng.QKV, eng.QKV, eng.QKV,
                                    o=o, lse=lse), lab)

# training attention backward (tcgen05): CTA (key tile 0, batch row 0): 16 iterations, first 6 stamped
dout = torch.randn(Bm * T, eng.QD, device=dev, generator=gen).to(torch.bfloat16)
dqkv = torch.zeros(Bm * T, eng.QKV, dtype=torch.bfloat16, device=dev)
pos_t = torch.arange(T, dtype=torch.int32, device=dev)
lab = {0: "start (after setup)"}
for it in range(6):
    lab[1 + 5 * it] = f"it{it}: S/dP ready"
    lab[2 + 5 * it] = f"it{it}: P/dS computed (regs)"
    lab[3 + 5 * it] = f"it{it}: MMA2(it-1) done"
    lab[4 + 5 * it] = f"it{it}: P/dS in smem, signalled"
    lab[5 + 5 * it] = f"it{it}: dQ(it-1) drained"
timeline("attn_bwd_tc (key tile 0)",
         lambda: ops.attn_train_bwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], o, dout, lse, pos_t, eng.timescale, Bm, T, eng.Nq, eng.Nkv,
                                    eng.D, eng.QKV, eng.QKV, eng.QKV, dqkv, dqkv[:, QD:], dqkv[:, QD + KD:], eng.QKV, eng.QKV,
                                    eng.QKV), lab)
act = torch.empty(B, dtype=torch.int32, device=dev)
lp = torch.empty(B, dtype=torch.float32, device=dev)
ao = torch.randn(B, eng.H, device=dev, generator=gen).to(torch.bfloat16)
glu_labels = {0: "start", 1: "setup done (barriers, tmem, cluster sync)", 2: "x1 accumulator ready", 3: "x1 + norm done", 4: "ug0 ready",
              5: "h0 written", 6: "ug1 ready", 7: "h1 written", 8: "ug2 ready", 9: "h2 written", 20: "D ready", 21: "partial pushed",
              22: "all partials in", 23: "output stored (+ policy head)", 24: "exit sync", 25: "head: rows summed, block output stored",
              26: "head: output norm done", 27: "head: logits reduced", 28: "head: noise + arg-max done"}
timeline("glu_fused + out projection", lambda: ops.glu_fused(ao, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F,
                                                             norm_scale=p[pre + "ffn_norm.scale"], wt_o=w[pre + "wt_o"]), glu_labels)
timeline("glu_fused_head", lambda: ops.glu_fused_head(ao, w[pre + "wt_ug64"], w[pre + "wt_d"], x, x2, eng.F, p[pre + "ffn_norm.scale"],
                                                      w[pre + "wt_o"], p["output_norm.scale"], p["action_embedder.embedding_table"],
                                                      act, lp, seed=1, stream_id=2), glu_labels)
