#!/usr/bin/env python
"""Warm timing of one optimizer step (Muon partition / AdamW) + weight re-staging on the parameter set of the named
configs (default return_2_layers):  python scripts/microbench_optimizer.py [config ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxrl_b200.config import NAMED  # noqa: E402
from jaxrl_b200.trainer import PPOTrainer  # noqa: E402
from jaxrl_b200 import ops  # noqa: E402
from dataclasses import replace  # noqa: E402

for name in (sys.argv[1:] or ["return_2_layers"]):
    run = replace(NAMED[name], num_agents=256)
    tr = PPOTrainer(run, use_graphs=False)
    eng, o, opt = tr.engine, tr.o, tr.opt
    eng.p.grad.normal_(std=1e-3)
    m = o.muon


    def step():
        ops.muon_step(eng.p.flat, eng.p.grad, o.mu, o.nu, o.scratch, eng.p.n_adam, m.table, m.nmat, m.x_total, m.a_total, m.max_r,
                      m.max_c, m.ns_ws, o.step, opt.learning_rate, 1000, opt.beta1, opt.beta2, 1e-8, opt.weight_decay, opt.max_norm,
                      grad_scale=1.0, norm_out=o.norm, ws=m.ws)
        eng.stage_weights(rollout=False)


    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                step()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: muon step + weight staging: {e0.elapsed_time(e1) * 100:.1f} us per optimizer step ({m.nmat} Muon matrices)")
