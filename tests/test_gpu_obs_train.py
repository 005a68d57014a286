"""GPU parity tests of the fused training-side observation encoder (jaxrl_b200/csrc/obs_train.cu) against the CPU
oracle's GridCnnObsEncoder restatement (reference mapox_trainer/model/observation.py:84-96, network.py:303-309) and its
autograd (what nnx.grad(ppo_loss) derives, train.py:236-238)."""

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402
from _util import assert_close_bf16  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from jaxrl_b200 import ops as _ops
    return _ops


def g(seed):
    return torch.Generator().manual_seed(seed)


def _params(gen, C, H, A, tasks):
    p = {"obs_encoder.layers.0.kernel": torch.randn(3, 3, C, 16, generator=gen) * 0.3,
         "obs_encoder.layers.0.bias": torch.randn(16, generator=gen) * 0.1,
         "obs_encoder.layers.1.kernel": torch.randn(3, 3, 16, 32, generator=gen) / 12.0,
         "obs_encoder.layers.1.bias": torch.randn(32, generator=gen) * 0.1,
         "obs_encoder.layers.2.kernel": torch.randn(3, 3, 32, H, generator=gen) / 17.0,
         "obs_encoder.layers.2.bias": torch.randn(H, generator=gen) * 0.1,
         "reward_encoder.kernel": torch.randn(1, H, generator=gen), "reward_encoder.bias": torch.randn(H, generator=gen) * 0.1,
         "action_embedder.embedding_table": torch.randn(A, H, generator=gen) * 0.1}
    if tasks:
        p["task_embedder.embedding_table"] = torch.randn(tasks, H, generator=gen) * 0.1
    return p


def _oracle_stages(p, cfg, obs):
    """z1, a1, z2, a2, obs_emb of the oracle encoder (bf16 at every rounding point)."""
    x = O.concat_one_hot(obs, cfg.obs_max_value, BF16)
    z1 = O.conv_valid(x, p["obs_encoder.layers.0.kernel"], p["obs_encoder.layers.0.bias"], cfg.cnn_strides[0], BF16)
    a1 = O.gelu_tanh(z1)
    z2 = O.conv_valid(a1, p["obs_encoder.layers.1.kernel"], p["obs_encoder.layers.1.bias"], cfg.cnn_strides[1], BF16)
    a2 = O.gelu_tanh(z2)
    z3 = O.conv_valid(a2, p["obs_encoder.layers.2.kernel"], p["obs_encoder.layers.2.bias"], cfg.cnn_strides[2], BF16)
    return z1, a1, z2, a2, z3.reshape(z3.shape[0], -1)


def _image(ops, p, sizes, C, H):
    b = lambda t: t.to(BF16).to(DEV)
    w1 = b(p["obs_encoder.layers.0.kernel"].reshape(9 * C, 16))
    wt2 = b(p["obs_encoder.layers.1.kernel"].reshape(144, 32).t().contiguous())
    wt3 = b(p["obs_encoder.layers.2.kernel"].reshape(288, H).t().contiguous())
    return ops.obs_fused_prep(sizes, w1, b(p["obs_encoder.layers.0.bias"]), wt2, b(p["obs_encoder.layers.1.bias"]), wt3,
                              b(p["obs_encoder.layers.2.bias"]), ops.obs_fused_image(DEV))


def _run_fwd(ops, p, sizes, obs, reward, la, tk, tasks, M, H=128, emb=None):
    d = lambda t: t.to(DEV)
    buf = ops.obs_train_buffers(M, DEV)
    for k in ("z1t", "a1t", "z2t", "dz2t"):
        buf[k].fill_(float("nan"))
    a2 = torch.full((M, 288), float("nan"), dtype=BF16, device=DEV)
    out = torch.full((M, H), float("nan"), dtype=BF16, device=DEV)
    image = _image(ops, p, sizes, sum(sizes), H)
    ops.obs_train_fwd(d(obs), sizes, image, d(reward), d(la), None if tk is None else d(tk),
                      d(p["reward_encoder.kernel"].reshape(-1)), d(p["reward_encoder.bias"]),
                      d(p["action_embedder.embedding_table"]), d(p["task_embedder.embedding_table"]) if tasks else None,
                      buf["z1t"], buf["a1t"], buf["z2t"], a2, out, M, 11, 11, obs_emb=emb)
    return buf, a2, out, image


CASES = [(128, (5, 5), 0), (1, (5, 5), 0), (77, (3, 2, 2), 3), (300, (5, 5), 0), (128 * 150 + 5, (5, 5), 0), (200, (6,), 0),
         (130, (2, 2, 1, 1), 2)]


@pytest.mark.parametrize("M,sizes,tasks", CASES)
def test_obs_train_fwd(ops, M, sizes, tasks):
    """x and every saved activation (z1, a1, z2, a2) of the one-kernel training forward vs the oracle stages; ragged
    tails, several tiles per CTA (M > 148 * 128), 1 to 4 observation components, out-of-range cell values."""
    from jaxrl_b200.config import ModelConfig
    gen = g(M + len(sizes) + tasks)
    K, C, H, A = len(sizes), sum(sizes), 128, 6
    assert ops.obs_train_supported(sizes)
    cfg = ModelConfig(obs_shape=(11, 11, K), obs_max_value=tuple(sizes), action_dim=A, task_count=max(tasks, 1))
    comps = [torch.randint(0, n + 1, (M, 11, 11), generator=gen) for n in sizes]      # n = out of range: all-zero one-hot
    obs = torch.stack(comps, -1).to(torch.uint8)
    p = _params(gen, C, H, A, tasks)
    reward = torch.randn(M, generator=gen)
    la = torch.randint(-1, A, (M,), generator=gen).to(torch.int32)
    tk = torch.randint(0, tasks, (M,), generator=gen).to(torch.int32) if tasks else None
    z1r, a1r, z2r, a2r, embr = _oracle_stages(p, cfg, obs)
    ts = O.TimeStep(obs=obs[:, None], time=torch.zeros(M, 1, dtype=torch.int32), terminated=torch.zeros(M, 1, dtype=torch.bool),
                    last_action=la[:, None], reward=reward[:, None], action_mask=None, task_ids=None if tk is None else tk[:, None])
    x_ref = O.model_embed(p, cfg, ts)[:, 0]
    emb = torch.full((M, H), float("nan"), dtype=BF16, device=DEV)
    buf, a2, out, _ = _run_fwd(ops, p, sizes, obs, reward, la, tk, tasks, M, emb=emb)
    torch.cuda.synchronize()
    kw = dict(rtol=1e-2, abs_floor_frac=1e-2, allow_frac=2e-3, max_err_frac=0.1)
    z1, a1, z2 = (ops.obs_tile_to_rows(buf[k], M) for k in ("z1t", "a1t", "z2t"))
    assert_close_bf16(z1, z1r.reshape(M, 25, 16), f"obs_train_fwd z1 M={M}", rtol=1e-2, abs_floor_frac=4e-3)
    assert_close_bf16(a1, a1r.reshape(M, 25, 16), f"obs_train_fwd a1 M={M}", rtol=1e-2, abs_floor_frac=4e-3)
    assert_close_bf16(z2, z2r.reshape(M, 9, 32), f"obs_train_fwd z2 M={M}", **kw)
    assert_close_bf16(a2, a2r.reshape(M, 288), f"obs_train_fwd a2 M={M}", **kw)
    assert_close_bf16(emb, embr, f"obs_train_fwd encoder output M={M}", **kw)
    assert_close_bf16(out, x_ref, f"obs_train_fwd x M={M}", **kw)


def test_obs_train_fwd_matches_rollout_kernel(ops):
    """The training forward and the rollout-step kernel (obs_fused.cu) share every rounding point: identical bits."""
    M, sizes, A, H = 1000, (5, 5), 8, 128
    gen = g(11)
    C = sum(sizes)
    obs = torch.stack([torch.randint(0, n, (M, 11, 11), generator=gen) for n in sizes], -1).to(torch.uint8)
    p = _params(gen, C, H, A, 0)
    reward = torch.randn(M, generator=gen)
    la = torch.randint(-1, A, (M,), generator=gen).to(torch.int32)
    d = lambda t: t.to(DEV)
    _, _, x_train, image = _run_fwd(ops, p, sizes, obs, reward, la, None, 0, M)
    x_roll = torch.empty(M, H, dtype=BF16, device=DEV)
    ops.obs_embed_fused(d(obs), sizes, image, d(reward), d(la), None, d(p["reward_encoder.kernel"].reshape(-1)),
                        d(p["reward_encoder.bias"]), d(p["action_embedder.embedding_table"]), None, x_roll, M, 11, 11)
    torch.cuda.synchronize()
    assert torch.equal(x_roll, x_train)


@pytest.mark.parametrize("M,sizes", [(128, (5, 5)), (1, (5, 5)), (77, (3, 2, 2)), (300, (5, 5)), (128 * 150 + 5, (5, 5)), (200, (6,))])
def test_obs_train_bwd(ops, M, sizes):
    """Data gradients (dz2, dz1), conv1 / conv2 bias gradients and the tap-wise conv2 weight gradient of the fused
    backward vs autograd through the oracle encoder (what nnx.grad derives, train.py:236-238)."""
    from jaxrl_b200.config import ModelConfig
    gen = g(1000 + M + len(sizes))
    K, C, H, A = len(sizes), sum(sizes), 128, 6
    cfg = ModelConfig(obs_shape=(11, 11, K), obs_max_value=tuple(sizes), action_dim=A)
    obs = torch.stack([torch.randint(0, n, (M, 11, 11), generator=gen) for n in sizes], -1).to(torch.uint8)
    p = _params(gen, C, H, A, 0)
    reward = torch.randn(M, generator=gen)
    la = torch.randint(-1, A, (M,), generator=gen).to(torch.int32)
    dx0 = (torch.randn(M, H, generator=gen) * 0.05).to(BF16)
    # oracle: autograd through the staged encoder with the cotangent dx0 on its output
    pg = {k: v.clone().requires_grad_(True) for k, v in p.items() if k.startswith("obs_encoder")}
    x = O.concat_one_hot(obs, cfg.obs_max_value, BF16)
    z1 = O.conv_valid(x, pg["obs_encoder.layers.0.kernel"], pg["obs_encoder.layers.0.bias"], cfg.cnn_strides[0], BF16)
    z1.retain_grad()
    z2 = O.conv_valid(O.gelu_tanh(z1), pg["obs_encoder.layers.1.kernel"], pg["obs_encoder.layers.1.bias"], cfg.cnn_strides[1], BF16)
    z2.retain_grad()
    z3 = O.conv_valid(O.gelu_tanh(z2), pg["obs_encoder.layers.2.kernel"], pg["obs_encoder.layers.2.bias"], cfg.cnn_strides[2], BF16)
    (z3.reshape(M, H).float() * dx0.float()).sum().backward()
    # kernels
    buf, a2, _, image = _run_fwd(ops, p, sizes, obs, reward, la, None, 0, M)
    dz1 = torch.full((M * 25, 16), float("nan"), dtype=BF16, device=DEV)
    db1 = torch.full((16,), 0.5, device=DEV)
    db2 = torch.full((32,), -0.25, device=DEV)
    dw2 = torch.full((3, 3, 16, 32), 0.125, device=DEV)
    ops.obs_train_bwd(dx0.to(DEV), buf["z1t"], buf["z2t"], image, buf["dz2t"], dz1, db1, db2, M, buf["ws"])
    ops.obs_conv2_wgrad(buf["a1t"], buf["dz2t"], dw2, M, buf["ws"])
    torch.cuda.synchronize()
    kw = dict(rtol=2e-2, abs_floor_frac=2e-2, allow_frac=2e-3, max_err_frac=0.2)
    assert_close_bf16(ops.obs_tile_to_rows(buf["dz2t"], M), z2.grad.reshape(M, 9, 32), f"obs_train_bwd dz2 M={M}", **kw)
    assert_close_bf16(dz1.view(M, 25, 16), z1.grad.reshape(M, 25, 16), f"obs_train_bwd dz1 M={M}", **kw)

    def rel(got, ref):
        return (got.float().cpu() - ref).norm().item() / (ref.norm().item() + 1e-12)

    assert rel(db1 - 0.5, pg["obs_encoder.layers.0.bias"].grad) < 1e-2, "db1"          # += semantics
    assert rel(db2 + 0.25, pg["obs_encoder.layers.1.bias"].grad) < 1e-2, "db2"
    assert rel(dw2 - 0.125, pg["obs_encoder.layers.1.kernel"].grad) < 1e-2, "dW2"
    # determinism: fixed summation orders everywhere
    dz1b, db1b, db2b, dw2b = torch.empty_like(dz1), torch.full_like(db1, 0.5), torch.full_like(db2, -0.25), torch.full_like(dw2, 0.125)
    ops.obs_train_bwd(dx0.to(DEV), buf["z1t"], buf["z2t"], image, buf["dz2t"], dz1b, db1b, db2b, M, buf["ws"])
    ops.obs_conv2_wgrad(buf["a1t"], buf["dz2t"], dw2b, M, buf["ws"])
    torch.cuda.synchronize()
    assert torch.equal(dz1, dz1b) and torch.equal(db1, db1b) and torch.equal(db2, db2b) and torch.equal(dw2, dw2b)


def test_engine_encoder_gradients_fused_vs_unfused():
    """PolicyEngine._embed_fwd / _embed_bwd with the fused training encoder against the im2col + GEMM chain on the same
    parameters: every obs_encoder gradient agrees to bf16 accuracy."""
    from jaxrl_b200.config import ModelConfig
    from jaxrl_b200.engine import PolicyEngine
    cfg = ModelConfig(num_layers=1, max_seq_length=64, action_dim=6)
    Bm, T = 6, 64
    M = Bm * T
    gen = g(5)
    obs = torch.stack([torch.randint(0, n, (M, 11, 11), generator=gen) for n in cfg.obs_max_value], -1).to(torch.uint8).to(DEV)
    reward = torch.randn(M, generator=gen).to(DEV)
    la = torch.randint(0, cfg.action_dim, (M,), generator=gen).to(torch.int32).to(DEV)
    dx0 = (torch.randn(M, cfg.hidden_features, generator=gen) * 0.05).to(BF16).to(DEV)
    res = {}
    for fused in (False, True):
        eng = PolicyEngine(cfg, DEV, seed=3, bias_noise=0.05, fused_obs_train=fused)
        assert eng.fused_obs_train == fused
        ws = eng.alloc_train(Bm, T)
        eng._embed_fwd(ws, M, obs, reward, la, None, ws.x0, train=True)
        eng.p.grad.zero_()
        eng._embed_bwd(ws, M, dx0.clone(), reward, la, None, obs)
        torch.cuda.synchronize()
        res[fused] = (ws.x0.clone(), {k: v.clone() for k, v in eng.p.grad_dict().items() if k.startswith("obs_encoder")})
    assert torch.equal(res[False][0], res[True][0])
    for k, ref in res[False][1].items():
        got = res[True][1][k]
        r = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
        assert r < 1e-2, f"{k}: rel l2 {r}"


# ---------------------------------------------------------------------------------------------- value path backward
@pytest.mark.parametrize("M,Vh,NL", [(128, 768, 51), (1, 64, 64), (300, 768, 51), (128 * 150 + 37, 768, 51), (77, 128, 17),
                                     (128 * 297 + 5, 192, 51)])
def test_value_bwd_fused(ops, M, Vh, NL):
    """One-kernel backward of the value MLP + fp32 head vs autograd through the oracle (network.py:335-341, values.py:34,
    40-41): pv, dz, d xf, the bias gradient, and bit-identity with the GEMM + gelu_bwd chain it replaces."""
    import math
    gen = g(M + Vh + NL)
    H = 128
    xf = torch.randn(M, H, generator=gen).to(BF16)
    wvm = torch.randn(H, Vh, generator=gen) / math.sqrt(H)
    bvm = torch.randn(Vh, generator=gen) * 0.1
    wvh = torch.randn(Vh, NL, generator=gen) / math.sqrt(Vh)
    dvl = torch.zeros(M, 64, dtype=BF16)
    dvl[:, :NL] = (torch.randn(M, NL, generator=gen) * 0.02).to(BF16)
    # oracle: z -> pv -> logits (fp32 head), cotangent dvl on the logits
    xr = xf.clone().requires_grad_(True)
    bvr = bvm.clone().requires_grad_(True)
    z = O.linear(xr, wvm, bvr)
    z.retain_grad()
    pv_ref = O.gelu_tanh(z)
    vl = O.linear(pv_ref, wvh, None, dtype=None)
    (vl * dvl[:, :NL].float()).sum().backward()
    # kernel
    d = lambda t: t.to(DEV)
    wt_vm = d(wvm.t().contiguous().to(BF16))
    w_vh_pad = torch.zeros(Vh, 64, dtype=BF16)
    w_vh_pad[:, :NL] = wvh.to(BF16)
    pv = torch.full((M, Vh), float("nan"), dtype=BF16, device=DEV)
    dz = torch.full((M, Vh), float("nan"), dtype=BF16, device=DEV)
    dxf = torch.full((M, H), float("nan"), dtype=BF16, device=DEV)
    db = torch.full((Vh,), 0.5, device=DEV)
    ops.value_bwd_fused(d(xf), d(dvl), wt_vm, d(bvm.to(BF16)), d(w_vh_pad), pv, dz, dxf, db)
    torch.cuda.synchronize()
    kw = dict(rtol=2e-2, abs_floor_frac=2e-2, allow_frac=2e-3, max_err_frac=0.2)
    assert_close_bf16(pv, pv_ref, f"value_bwd_fused pv M={M}", rtol=1e-2, abs_floor_frac=1e-2, allow_frac=2e-3, max_err_frac=0.1)
    assert_close_bf16(dz, z.grad, f"value_bwd_fused dz M={M}", **kw)
    assert_close_bf16(dxf, xr.grad, f"value_bwd_fused dxf M={M}", **kw)
    rel = ((db - 0.5).cpu() - bvr.grad).norm().item() / (bvr.grad.norm().item() + 1e-12)
    assert rel < 1e-2, f"value-MLP bias gradient rel l2 {rel}"
    # the chain it replaces: GEMM (+bias, gelu, pre-activation) / dpv GEMM / gelu_bwd / dxf GEMM - same rounding points
    pv2 = torch.empty(M, Vh, dtype=BF16, device=DEV)
    zv2 = torch.empty(M, Vh, dtype=BF16, device=DEV)
    dpv2 = torch.empty(M, Vh, dtype=BF16, device=DEV)
    dxf2 = torch.empty(M, H, dtype=BF16, device=DEV)
    ops.linear(d(xf), wt_vm, bias=d(bvm.to(BF16)), act=1, out=pv2, y_pre=zv2)
    ops.linear(d(dvl), d(w_vh_pad), out=dpv2)
    ops.gelu_bwd(dpv2, zv2, out=dpv2)
    ops.linear(dpv2, d(wvm.to(BF16)), out=dxf2)
    torch.cuda.synchronize()
    assert torch.equal(pv, pv2), "pv differs from the unfused chain"
    assert torch.equal(dz, dpv2), "dz differs from the unfused chain"
    assert_close_bf16(dxf, dxf2, f"value_bwd_fused dxf vs chain M={M}", rtol=1e-2, abs_floor_frac=4e-3)
    # deterministic
    db2 = torch.full((Vh,), 0.5, device=DEV)
    ops.value_bwd_fused(d(xf), d(dvl), wt_vm, d(bvm.to(BF16)), d(w_vh_pad), pv2, dpv2, dxf2, db2)
    torch.cuda.synchronize()
    assert torch.equal(db, db2) and torch.equal(dxf, dxf2)


def test_engine_value_path_fused_vs_unfused():
    """forward_train + loss_and_backward with the one-kernel value path (forward and backward) against the GEMM chain:
    losses and every gradient agree."""
    from jaxrl_b200.config import ModelConfig, PPOConfig
    from jaxrl_b200.engine import PolicyEngine
    cfg = ModelConfig(num_layers=1, max_seq_length=64, action_dim=6)
    hyp = PPOConfig()
    Bm, T = 5, 64
    M = Bm * T
    gen = g(9)
    obs = torch.stack([torch.randint(0, n, (Bm, T, 11, 11), generator=gen) for n in cfg.obs_max_value], -1).to(torch.uint8).to(DEV)
    batch = dict(last_rewards=torch.randn(M, generator=gen) * 0.1, last_actions=torch.randint(0, 6, (M,), generator=gen).to(torch.int32),
                 actions=torch.randint(0, 6, (M,), generator=gen).to(torch.int32), log_prob=-torch.rand(M, generator=gen) * 2 - 0.2,
                 advantages=torch.randn(M, generator=gen), targets=torch.randn(M, generator=gen) * 2)
    batch = {k: v.to(DEV) for k, v in batch.items()}
    batch["obs"] = obs
    res = {}
    for fused in (False, True):
        eng = PolicyEngine(cfg, DEV, seed=3, bias_noise=0.05, fused_value_train=fused)
        assert eng.fused_value_train == fused
        ws = eng.alloc_train(Bm, T)
        eng.forward_train(ws, obs, batch["last_rewards"], batch["last_actions"])
        eng.p.grad.zero_()
        losses = eng.loss_and_backward(ws, batch, hyp, progress=0.0).clone()
        torch.cuda.synchronize()
        res[fused] = (losses.cpu(), ws.vlogits.clone(), eng.p.grad_dict())
    torch.testing.assert_close(res[True][0][:3], res[False][0][:3], rtol=1e-4, atol=1e-5)     # [3] is a spare slot
    torch.testing.assert_close(res[True][1], res[False][1], rtol=1e-3, atol=1e-3)
    for k, ref in res[False][2].items():
        got = res[True][2][k]
        r = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
        assert r < 2e-2, f"{k}: rel l2 {r}"
