"""world_size-2 gloo tests (CPU) of the data-parallel host logic: env sharding + the two collectives give exactly
the single-device result (global advantage normalisation, mean gradient)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxrl_b200 import dist as mdist
from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, T = 8, 16
        rng = np.random.default_rng(0)
        r, v = rng.standard_normal((B, T)).astype(np.float32), rng.standard_normal((B, T + 1)).astype(np.float32)
        term = rng.random((B, T)) < 0.1
        lo, hi = mdist.shard_bounds(B, rank, world)
        adv, tgt = O.calculate_advantage(r[lo:hi], v[lo:hi], term[lo:hi], 0.97, 0.9, False)
        stats = torch.tensor([adv.astype(np.float64).sum(), (adv.astype(np.float64) ** 2).sum()], dtype=torch.float64)
        count = mdist.allreduce_adv_moments(stats, adv.size)
        mean = stats[0].item() / count
        std = (stats[1].item() / count - mean * mean) ** 0.5
        adv_n = (adv - np.float32(mean)) / (np.float32(std) + np.float32(1e-8))
        # gradient: local mean-loss gradients, sum all-reduce, scale 1/world == gradient of the global mean
        w = torch.arange(4, dtype=torch.float32)
        x = torch.from_numpy(r[lo:hi, :4])
        g_local = (2 * (x @ w)[:, None] * x).mean(0)           # d/dw mean((x.w)^2) over the shard
        g = g_local.clone()
        scale = mdist.allreduce_gradients(g)
        # the trainer reduces the flat gradient as two partitions (Muon matrices early, AdamW leaves late): same result
        g2 = g_local.clone()
        n_adam = 1
        s_hi = mdist.allreduce_gradients(g2[n_adam:])
        s_lo = mdist.allreduce_gradients(g2[:n_adam])
        assert s_hi == s_lo == scale and torch.equal(g2, g)
        p = torch.zeros(3) if rank else torch.ones(3)
        mdist.broadcast_params(p)
        q.put((rank, lo, hi, adv_n, (g * scale).numpy(), p.numpy(), count))
    finally:
        dist.destroy_process_group()


def test_env_sharded_collectives_match_single_device():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, q)) for rk in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort(key=lambda t: t[0])
    B, T = 8, 16
    rng = np.random.default_rng(0)
    r, v = rng.standard_normal((B, T)).astype(np.float32), rng.standard_normal((B, T + 1)).astype(np.float32)
    term = rng.random((B, T)) < 0.1
    ref_adv, _ = O.calculate_advantage(r, v, term, 0.97, 0.9, True)
    got = np.concatenate([o[3] for o in out], axis=0)
    np.testing.assert_allclose(got, ref_adv, rtol=1e-5, atol=1e-5)
    w = torch.arange(4, dtype=torch.float32)
    x = torch.from_numpy(r[:, :4])
    g_ref = (2 * (x @ w)[:, None] * x).mean(0).numpy()
    for o in out:
        np.testing.assert_allclose(o[4], g_ref, rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(o[5], np.ones(3, np.float32))       # rank-0 weights everywhere
        assert o[6] == B * T
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 4, 4, 8)


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jaxrl_b200.config import ModelConfig
        from jaxrl_b200.params import ParamStore, is_muon_param
        cfg = ModelConfig(num_layers=2, max_seq_length=16)
        ps = ParamStore(cfg, "cpu", seed=5, world=world)
        n_adam, shard = ps.n_adam, ps.muon_shard
        # every rank's gradient of every Muon matrix = (rank + 1) * (index of the matrix + 1)
        names = [n for n, sh, _ in ps.specs if is_muon_param(n, sh)]
        for i, n in enumerate(names):
            ps.gviews[n].fill_((rank + 1) * (i + 1))
        scale = mdist.reduce_scatter_shard(ps.grad[n_adam:], shard)
        total = sum(r + 1 for r in range(world))
        mine = [n for n in names if ps.muon_owner[n] == rank]
        ok_rs = all(torch.equal(ps.gviews[n], torch.full_like(ps.gviews[n], total * (names.index(n) + 1))) for n in mine)
        # "update" of the owned matrices = -index, then all-gather: every rank must hold every matrix's update
        upd = torch.zeros_like(ps.flat)
        for n in mine:
            off, shape = ps.offsets[n]
            upd[off:off + ps.views[n].numel()] = -(names.index(n) + 1)
        mdist.all_gather_shards(upd[n_adam:], shard)
        ok_ag = all((upd[ps.offsets[n][0]:ps.offsets[n][0] + ps.views[n].numel()] == -(i + 1)).all().item()
                    for i, n in enumerate(names))
        q.put((rank, scale, ok_rs, ok_ag, len(mine), shard, ps.numel))
    finally:
        dist.destroy_process_group()


def test_sharded_muon_partition_collectives():
    """reduce-scatter / all-gather of the per-rank Muon shards (params.ParamStore(world=2)) under gloo."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(rk, world, port, q)) for rk in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scale, ok_rs, ok_ag, nmine, shard, numel in out:
        assert scale == 0.5 and ok_rs and ok_ag and nmine >= 1
    assert out[0][5] == out[1][5] and out[0][6] == out[1][6]


def test_muon_owner_assignment_is_balanced_and_layout_preserves_values():
    from jaxrl_b200.config import NAMED
    from jaxrl_b200.params import ParamStore, assign_muon_owners, is_muon_param
    assert assign_muon_owners([10, 10, 10, 1], 2) == [0, 1, 0, 1]
    cfg = NAMED["return_2_layers"].model
    a = ParamStore(cfg, "cpu", seed=3)
    for world in (2, 4, 8):
        b = ParamStore(cfg, "cpu", seed=3, world=world)
        assert all(torch.equal(a.views[n], b.views[n]) for n in a.views)          # values do not depend on the layout
        sizes = [0] * world
        for n, sh, _ in b.specs:
            if is_muon_param(n, sh):
                r = b.muon_owner[n]
                off = b.offsets[n][0]
                assert b.n_adam + r * b.muon_shard <= off and off + a.views[n].numel() <= b.n_adam + (r + 1) * b.muon_shard
                sizes[r] += a.views[n].numel()
        assert max(sizes) - min(sizes) <= 98304          # at most one ffn matrix of imbalance
        nm = [b.muon_table(r)[1] for r in range(world)]
        assert sum(nm) == a.muon_table()[1]


def test_shard_bounds_rejects_ragged():
    with pytest.raises(ValueError):
        mdist.shard_bounds(10, 0, 4)
