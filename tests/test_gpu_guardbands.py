"""Out-of-bounds write detection without compute-sanitizer (closed on this pool: profiles/r02_compute_sanitizer_closed.log).

Every output and in-place buffer of the fused hot-path kernels is carved out of a larger allocation filled with a canary
pattern; after the launch the guard bands on both sides must be untouched.  Shapes are deliberately ragged (row counts
that are not multiples of the tile sizes) so that the kernels' tail predication is what is being exercised."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
BF16 = torch.bfloat16
GUARD = 4096          # bytes on each side


class Guarded:
    """A tensor of `shape` / `dtype` in the middle of a canary-filled byte buffer."""

    def __init__(self, shape, dtype, fill=None):
        n = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((n + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
        self.t = self.raw[GUARD:GUARD + n].view(dtype).view(shape)
        if fill is not None:
            self.t.copy_(fill.to(DEV))
        self.n = n

    def check(self, name):
        torch.cuda.synchronize()
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.n:]
        assert bool((lo == 0xA5).all()) and bool((hi == 0xA5).all()), f"{name}: write outside the buffer"


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("B,step,S", [(1, 0, 16), (13, 5, 16), (15, 40, 64), (29, 700, 512), (2075, 300, 512)])
@pytest.mark.parametrize("H", [128, 256])
def test_attn_decode_fused_guard(B, step, S, H):
    from jaxrl_b200 import ops
    from _util import rope_timescale
    gen = _g(B + step + H)
    D, Nq, Nkv = 32, 4, 1
    x = Guarded((B, H), BF16, (torch.randn(B, H, generator=gen)).to(BF16))
    kc = Guarded((B, S, Nkv, D), BF16, torch.randn(B, S, Nkv, D, generator=gen).to(BF16))
    vc = Guarded((B, S, Nkv, D), BF16, torch.randn(B, S, Nkv, D, generator=gen).to(BF16))
    out = Guarded((B, Nq * D), BF16)
    w = (torch.randn((Nq + 2 * Nkv) * D, H, generator=gen) / math.sqrt(H)).to(BF16).to(DEV)
    pos = torch.full((B,), step, dtype=torch.int32, device=DEV)
    ops.attn_decode_fused(x.t, torch.ones(H, device=DEV), w, pos, rope_timescale(D).to(DEV), kc.t, vc.t, Nq, Nkv, D, out=out.t)
    for name, b in (("x", x), ("kcache", kc), ("vcache", vc), ("out", out)):
        b.check(f"attn_decode_fused {name} B={B}")
    assert torch.isfinite(out.t.float()).all()


@pytest.mark.parametrize("M", [1, 31, 129, 1000, 4097])
def test_glu_fused_and_value_head_guard(M):
    from jaxrl_b200 import ops
    gen = _g(M)
    H, F, Vh, NL = 128, 768, 768, 51
    ao = Guarded((M, H), BF16, torch.randn(M, H, generator=gen).to(BF16))
    res = Guarded((M, H), BF16, torch.randn(M, H, generator=gen).to(BF16))
    out = Guarded((M, H), BF16)
    wu, wg = torch.randn(H, F, generator=gen) / 11.3, torch.randn(H, F, generator=gen) / 11.3
    wt_ug64 = ops.pack_glu_weight64(wu, wg).to(DEV)
    wt_d = (torch.randn(H, F, generator=gen) / 27.7).to(BF16).to(DEV)
    wt_o = (torch.randn(H, H, generator=gen) / 11.3).to(BF16).to(DEV)
    ops.glu_fused(ao.t, wt_ug64, wt_d, res.t, out.t, F, norm_scale=torch.ones(H, device=DEV), wt_o=wt_o)
    for name, b in (("ao", ao), ("residual", res), ("out", out)):
        b.check(f"glu_fused {name} M={M}")
    # training forward variant (h_out) and the fused backward
    h = Guarded((M, F), BF16)
    out2 = Guarded((M, H), BF16)
    ops.glu_fused(ao.t, wt_ug64, wt_d, res.t, out2.t, F, h_out=h.t)
    dug, dxn = Guarded((M, 2 * F), BF16), Guarded((M, H), BF16)
    ops.glu_bwd_fused(ao.t, res.t, wt_ug64, wt_d, F, dug=dug.t, dxn=dxn.t)
    for name, b in (("h_out", h), ("out (h_out variant)", out2), ("dug", dug), ("dxn", dxn)):
        b.check(f"glu_fused/bwd {name} M={M}")
    # value head
    value, vlog = Guarded((M,), torch.float32), Guarded((M, NL), torch.float32)
    wt_vm = (torch.randn(Vh, H, generator=gen) / 11.3).to(BF16).to(DEV)
    wvh = torch.randn(Vh, NL, generator=gen) / 27.7
    hi = wvh.t().to(BF16)
    hilo = torch.zeros(128, Vh, dtype=BF16)
    hilo[:NL] = hi
    hilo[64:64 + NL] = (wvh.t() - hi.float()).to(BF16)
    ops.value_head_fused(ao.t, wt_vm, torch.zeros(Vh, dtype=BF16, device=DEV), hilo.to(DEV), torch.zeros(NL, device=DEV),
                         torch.linspace(-10, 10, NL, device=DEV), value.t, vlogits=vlog.t, norm_scale=torch.ones(H, device=DEV))
    value.check(f"value_head_fused value M={M}")
    vlog.check(f"value_head_fused vlogits M={M}")


@pytest.mark.parametrize("M", [1, 33, 200, 4097])
def test_obs_embed_and_policy_head_guard(M):
    from jaxrl_b200 import ops
    from jaxrl_b200.config import ModelConfig
    from jaxrl_b200.engine import PolicyEngine
    cfg = ModelConfig(num_layers=1, max_seq_length=16, action_dim=6)
    eng = PolicyEngine(cfg, DEV, seed=1, bias_noise=0.05)
    gen = _g(M)
    obs = torch.stack([torch.randint(0, n, (M, 11, 11), generator=gen) for n in cfg.obs_max_value], -1).to(torch.uint8)
    og = Guarded(tuple(obs.shape), torch.uint8, obs)
    x = Guarded((M, 128), BF16)
    p = eng.p
    ops.obs_embed_fused(og.t, cfg.obs_max_value, eng.obs_image, torch.randn(M, generator=gen).to(DEV),
                        torch.randint(0, 6, (M,), generator=gen).to(torch.int32).to(DEV), None, p["reward_encoder.kernel"],
                        p["reward_encoder.bias"], p["action_embedder.embedding_table"], None, x.t, M, 11, 11)
    og.check(f"obs_embed_fused obs M={M}")
    x.check(f"obs_embed_fused x M={M}")
    act, lp, lg = Guarded((M,), torch.int32), Guarded((M,), torch.float32), Guarded((M, 6), torch.float32)
    ops.policy_head_sample(x.t, p["action_embedder.embedding_table"], act.t, lp.t, seed=3, stream_id=1,
                           norm_scale=p["output_norm.scale"], logits_out=lg.t)
    for name, b in (("action", act), ("log_prob", lp), ("logits", lg), ("x", x)):
        b.check(f"policy_head_sample {name} M={M}")
    assert int(act.t.min()) >= 0 and int(act.t.max()) < 6


@pytest.mark.parametrize("B,T", [(1, 128), (3, 256), (2, 512)])
def test_attn_train_guard(B, T):
    from jaxrl_b200 import ops
    from _util import rope_timescale
    gen = _g(B * T)
    Nq, Nkv, D = 4, 1, 32
    QD, QKV = Nq * D, (Nq + 2 * Nkv) * D
    M = B * T
    qkv = Guarded((M, QKV), BF16, torch.randn(M, QKV, generator=gen).to(BF16))
    o, lse = Guarded((M, QD), BF16), Guarded((M, Nq), torch.float32)
    ops.attn_train_fwd(qkv.t, qkv.t[:, QD:], qkv.t[:, QD + D:], B, T, Nq, Nkv, D, QKV, QKV, QKV, o=o.t, lse=lse.t)
    dqkv = Guarded((M, QKV), BF16)
    delta, dq_acc = Guarded((M, Nq), torch.float32), Guarded((M, QD), torch.float32)
    dout = torch.randn(M, QD, generator=gen).to(BF16).to(DEV)
    ops.attn_train_bwd(qkv.t, qkv.t[:, QD:], qkv.t[:, QD + D:], o.t, dout, lse.t, torch.arange(T, dtype=torch.int32, device=DEV),
                       rope_timescale(D).to(DEV), B, T, Nq, Nkv, D, QKV, QKV, QKV, dqkv.t, dqkv.t[:, QD:], dqkv.t[:, QD + D:],
                       QKV, QKV, QKV, delta=delta.t, dq_acc=dq_acc.t, pos_len=T)
    for name, b in (("qkv", qkv), ("o", o), ("lse", lse), ("dqkv", dqkv), ("delta", delta), ("dq_acc", dq_acc)):
        b.check(f"attn_train {name} B={B} T={T}")
    assert torch.isfinite(dqkv.t.float()).all()
