"""Host-side sequencing of the policy hot path over the C-ABI kernels (no arithmetic happens in python/torch).

Mirrors, call for call, what the reference's TransformerActorCritic.__call__ (network.py:300-343) does in its two
modes - decode with a KV-cache carry (rollout, train.py:95) and full-window training (ppo_loss, train.py:182) - plus
the hand-written transpose of the training pass that nnx.grad derives in the reference (train.py:236-238).

All buffers are preallocated workspaces so a decode step / a minibatch step is a fixed launch sequence that can be
captured into a CUDA graph.
"""

from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .config import ModelConfig
from .params import ParamStore, conv_out

BF16 = torch.bfloat16
F32 = torch.float32
I32 = torch.int32


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class PolicyEngine:
    def __init__(self, cfg: ModelConfig, device="cuda", seed: int = 0, params: Optional[ParamStore] = None,
                 bias_noise: float = 0.0, fused_rollout: bool = True, world: int = 1, fused_obs_train: bool = True,
                 fused_value_train: bool = True):
        """fused_rollout=False runs the rollout step through the general (per-op) kernel chain even where the
        one-kernel-per-half-block kernels apply - the chain every shape without a fused kernel takes (tests);
        fused_obs_train=False does the same for the training-side observation encoder (im2col + GEMM chain),
        fused_value_train=False for the training-side value path (GEMM + gelu_bwd chain)."""
        if cfg.head_dim != 32:
            raise _lib.MpxError("head_dim must be 32 on the B200 path")
        if cfg.obs_type not in ("grid_cnn", "grid_flattened"):
            raise _lib.MpxError(f"obs_type {cfg.obs_type!r} is not on the B200 path (grid_cnn / grid_flattened)")
        if cfg.value_hidden_dim is None:
            raise _lib.MpxError("value_hidden_dim=None is not on the B200 path yet")
        self.cfg = cfg
        self.dev = torch.device(device)
        self.p = params if params is not None else ParamStore(cfg, self.dev, seed, bias_noise, world=world)
        H, D = cfg.hidden_features, cfg.head_dim
        self.H, self.F, self.Vh, self.NL, self.A = H, cfg.ffn_size, cfg.value_hidden_dim, cfg.value_n_logits, cfg.action_dim
        self.Nq, self.Nkv, self.D = cfg.num_heads, cfg.num_kv_heads, D
        self.QD, self.KD = self.Nq * D, self.Nkv * D
        self.QKV = self.QD + 2 * self.KD
        self.L = cfg.num_layers
        glu_ok = (H == 128 and cfg.ffn_size % 64 == 0)
        self.fused_glu = glu_ok and fused_rollout                  # decode path: one kernel per GLU block
        self.fused_attn_decode = (H in (128, 256) and self.Nq == 4 and self.Nkv == 1 and self.D == 32 and fused_rollout)
        # training path: the same forward kernel (keeping h) and a one-kernel backward that recomputes u, g, dh on chip
        self.fused_glu_train = glu_ok
        self.fused_value = (H == 128 and cfg.value_hidden_dim is not None and cfg.value_hidden_dim % 64 == 0
                            and cfg.value_n_logits <= 64 and fused_rollout)   # decode path: one kernel for the value path
        # training path: forward = the same kernel writing the logits, backward = one kernel (value_bwd_fused.cu)
        self.fused_value_train = bool(fused_value_train and H == 128 and cfg.value_hidden_dim % 64 == 0
                                      and cfg.value_hidden_dim <= 1024 and cfg.value_n_logits <= 64)
        self.timescale = (torch.tensor(cfg.rope_max_wavelength, dtype=F32)
                          ** (2.0 * torch.arange(0, D // 2, dtype=F32) / D)).to(self.dev)
        sup = torch.linspace(cfg.value_min, cfg.value_max, cfg.value_n_logits + 1, dtype=F32)
        self.supports = sup.to(self.dev)
        self.centers = ((sup[:-1] + sup[1:]) / 2).to(self.dev)
        # conv geometry (observation.py:63-82)
        self.C0 = sum(cfg.obs_max_value)
        self.conv = []
        self.flat = cfg.obs_type == "grid_flattened"            # FlattenedObsEncoder (observation.py:98-145)
        if self.flat:
            self.cells = cfg.obs_shape[0] * cfg.obs_shape[1]
            self.fk = 4 * self.cells                             # dense in_features (embed_features = 4)
            self.fkp = _round_up(self.fk, 8)
        elif cfg.obs_type != "grid_cnn":
            raise _lib.MpxError(f"obs_type {cfg.obs_type!r} is not on the B200 path")
        ih, iw, cin = cfg.obs_shape[0], cfg.obs_shape[1], self.C0
        chans = list(cfg.cnn_channels) + [H]
        for j, ((kh, kw), (sh, sw), ch) in enumerate(zip(cfg.cnn_kernels, cfg.cnn_strides, chans) if not self.flat else ()):
            oh, ow = conv_out(ih, kh, sh), conv_out(iw, kw, sw)
            kp = _round_up(kh * kw * cin, 8)
            self.conv.append(SimpleNamespace(j=j, ih=ih, iw=iw, cin=cin, kh=kh, kw=kw, sh=sh, sw=sw, oh=oh, ow=ow,
                                             cout=ch, k=kh * kw * cin, kp=kp))
            ih, iw, cin = oh, ow, ch
        if not self.flat and (ih != 1 or iw != 1):
            raise _lib.MpxError("grid_cnn encoder must reduce the view to 1x1 (flatten == hidden_features)")
        for c in self.conv[1:]:
            if c.cin % 8 != 0:
                raise _lib.MpxError("hidden conv channels must be multiples of 8")
        # rollout path: the standard geometry (11x11 view, 3x3/2 -> 16, 3x3/1 -> 32, 3x3/1 -> 128) runs as one kernel
        std_geom = (not self.flat and H == 128 and tuple(cfg.obs_shape[:2]) == (11, 11) and len(self.conv) == 3
                    and all((c.kh, c.kw) == (3, 3) for c in self.conv)
                    and [(c.sh, c.sw) for c in self.conv] == [(2, 2), (1, 1), (1, 1)]
                    and [c.cout for c in self.conv] == [16, 32, 128] and len(cfg.obs_max_value) <= 4)
        self.fused_obs = fused_rollout and std_geom and math.prod(n + 1 for n in cfg.obs_max_value) <= 48
        # training path: the same geometry as one forward kernel + one data-gradient kernel (obs_train.cu)
        c0 = self.conv[0] if self.conv else None
        self.onehot_wgrad = (c0 is not None and c0.k <= 128 and (c0.kh, c0.kw) == (3, 3) and c0.cout <= 64
                             and c0.cout % 8 == 0 and len(cfg.obs_max_value) <= 4)
        self.fused_obs_train = bool(fused_obs_train and std_geom and self.onehot_wgrad
                                    and ops.obs_train_supported(cfg.obs_max_value))
        self.obs_image = ops.obs_fused_image(self.dev) if (self.fused_obs or self.fused_obs_train) else None
        # weight-gradient GEMMs run on a side stream next to the data-gradient chain; a second one carries the early
        # optimizer phase (Muon partition) under the observation-encoder backward
        self.side = torch.cuda.Stream(device=self.dev)
        self.side2 = torch.cuda.Stream(device=self.dev)
        self._alloc_staged()
        self.stage_weights()

    # ------------------------------------------------------------------------------------------ staged bf16 weights
    def _alloc_staged(self):
        cfg, dev, p = self.cfg, self.dev, self.p
        H, F, Vh, NL, QD, KD, QKV = self.H, self.F, self.Vh, self.NL, self.QD, self.KD, self.QKV
        z = lambda *s: torch.zeros(*s, dtype=BF16, device=dev)
        self.w: Dict[str, torch.Tensor] = {}
        descs: List[tuple] = []

        def add(src, src_ld, dst, dst_ld, rows, cols, mode):
            descs.append((src.data_ptr(), src_ld, dst.data_ptr(), dst_ld, rows, cols, mode))

        for i in range(self.L):
            pre = f"layers.{i}."
            wq = p[pre + "history.query_proj.kernel"].view(H, QD)
            wk = p[pre + "history.key_proj.kernel"].view(H, KD)
            wv = p[pre + "history.value_proj.kernel"].view(H, KD)
            wo = p[pre + "history.out.kernel"].view(QD, H)
            wt_qkv, w_qkv = z(QKV, H), z(H, QKV)
            add(wq, QD, wt_qkv, H, H, QD, 1)
            add(wk, KD, wt_qkv[QD:], H, H, KD, 1)
            add(wv, KD, wt_qkv[QD + KD:], H, H, KD, 1)
            add(wq, QD, w_qkv, QKV, H, QD, 0)
            add(wk, KD, w_qkv[:, QD:], QKV, H, KD, 0)
            add(wv, KD, w_qkv[:, QD + KD:], QKV, H, KD, 0)
            wt_o, w_o = z(H, QD), z(QD, H)
            add(wo, H, wt_o, QD, QD, H, 1)
            add(wo, H, w_o, H, QD, H, 0)
            wu, wg, wd = p[pre + "ffn.up_proj.kernel"], p[pre + "ffn.up_gate.kernel"], p[pre + "ffn.down_proj.kernel"]
            wt_ug, w_ug = z(2 * F, H), z(H, 2 * F)
            for j in range(F // 128):     # 256-row tiles: [128 up rows | 128 gate rows]
                add(wu[:, 128 * j:], F, wt_ug[256 * j:], H, H, 128, 1)
                add(wg[:, 128 * j:], F, wt_ug[256 * j + 128:], H, H, 128, 1)
            add(wu, F, w_ug, 2 * F, H, F, 0)
            add(wg, F, w_ug[:, F:], 2 * F, H, F, 0)
            wt_ug64 = z(2 * F, H)                                  # 128-row tiles [64 up | 64 gate] (fused decode GLU)
            for j in range(F // 64):
                add(wu[:, 64 * j:], F, wt_ug64[128 * j:], H, H, 64, 1)
                add(wg[:, 64 * j:], F, wt_ug64[128 * j + 64:], H, H, 64, 1)
            self.w[pre + "wt_ug64"] = wt_ug64
            wt_d, w_d = z(H, F), z(F, H)
            add(wd, H, wt_d, F, F, H, 1)
            add(wd, H, w_d, H, F, H, 0)
            self.w.update({pre + "wt_qkv": wt_qkv, pre + "w_qkv": w_qkv, pre + "wt_o": wt_o, pre + "w_o": w_o,
                           pre + "wt_ug": wt_ug, pre + "w_ug": w_ug, pre + "wt_d": wt_d, pre + "w_d": w_d})
        # value mlp / head
        wvm = p["value_mlp.kernel"]
        wt_vm, w_vm, b_vm = z(Vh, H), z(H, Vh), z(Vh)
        add(wvm, Vh, wt_vm, H, H, Vh, 1)
        add(wvm, Vh, w_vm, Vh, H, Vh, 0)
        add(p["value_mlp.bias"].view(1, Vh), Vh, b_vm, Vh, 1, Vh, 0)
        wvh = p["value_head.dense.kernel"]                     # (Vh, NL) fp32
        if NL > 64:
            raise _lib.MpxError("value_n_logits > 64 is not supported by the hi/lo tensor-core value head")
        wt_vh_hilo, w_vh_pad = z(128, Vh), z(Vh, 64)
        add(wvh, NL, wt_vh_hilo, Vh, Vh, NL, 1)                 # hi = bf16(W^T)
        add(wvh, NL, wt_vh_hilo[64:], Vh, Vh, NL, 3)            # lo = bf16(W^T - hi)
        add(wvh, NL, w_vh_pad, 64, Vh, NL, 0)
        self.w.update({"wt_vm": wt_vm, "w_vm": w_vm, "b_vm": b_vm, "wt_vh_hilo": wt_vh_hilo, "w_vh_pad": w_vh_pad})
        # conv layers as GEMMs: kernel (kh,kw,cin,cout) == (k, cout) matrix
        for c in self.conv:
            wk_ = p[f"obs_encoder.layers.{c.j}.kernel"].view(c.k, c.cout)
            wt_c, w_c, b_c = z(c.cout, c.kp), z(c.k, c.cout), z(c.cout)
            add(wk_, c.cout, wt_c, c.kp, c.k, c.cout, 1)
            add(wk_, c.cout, w_c, c.cout, c.k, c.cout, 0)
            add(p[f"obs_encoder.layers.{c.j}.bias"].view(1, c.cout), c.cout, b_c, c.cout, 1, c.cout, 0)
            self.w.update({f"wt_c{c.j}": wt_c, f"w_c{c.j}": w_c, f"b_c{c.j}": b_c})
        if self.flat:                                           # dense half of FlattenedObsEncoder: (4*cells, H) kernel
            wfd = p["obs_encoder.dense.kernel"]
            wt_fd, w_fd, b_fd = z(H, self.fkp), z(self.fkp, H), z(H)   # K padded to a multiple of 8 with zeros
            add(wfd, H, wt_fd, self.fkp, self.fk, H, 1)
            add(wfd, H, w_fd, H, self.fk, H, 0)
            add(p["obs_encoder.dense.bias"].view(1, H), H, b_fd, H, 1, H, 0)
            self.w.update({"wt_fd": wt_fd, "w_fd": w_fd, "b_fd": b_fd})
        arr = (_lib.PrepDesc * len(descs))()
        for n, d in enumerate(descs):
            arr[n] = _lib.PrepDesc(*d, 0)
        raw = bytes(arr)
        self._desc_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self._ndesc = len(descs)

    def stage_weights(self, rollout: bool = True):
        """fp32 master params -> bf16 GEMM operand copies (one launch).  rollout=True also rebuilds the weight image of
        the fused observation encoder, which only the rollout path reads (so once per update is enough)."""
        ops.prep_weights(self._desc_dev, self._ndesc)
        if rollout or self.fused_obs_train:                   # the fused training encoder reads the same image
            self.stage_rollout_weights()

    def stage_rollout_weights(self):
        if self.obs_image is not None:
            w = self.w
            ops.obs_fused_prep(self.cfg.obs_max_value, w["w_c0"], w["b_c0"], w["wt_c1"], w["b_c1"], w["wt_c2"], w["b_c2"],
                               self.obs_image)

    # ------------------------------------------------------------------------------------------ workspaces
    def _alloc_embed(self, ws, M: int, train: bool):
        dev = self.dev
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)
        n = len(self.conv)
        ws.cx, ws.cz, ws.ca = [], [], []
        if self.flat:
            ws.fe = e(M, self.fkp)                                # flattened cell embeddings (dense input)
            ws.oe = e(M, self.H)                                  # encoder output
            ws.dfe = e(M, self.fkp) if train else None
        fused_t = train and self.fused_obs_train
        if fused_t:
            ws.obs_t = ops.obs_train_buffers(M, dev)              # z1t, a1t, z2t, dz2t (tile layouts) + workspace
        for c in self.conv:
            rows = M * c.oh * c.ow
            # the last conv covers the whole remaining map: its im2col is the flattened previous activation;
            # the first conv reads the uint8 observation directly (its one-hot im2col is only built for the
            # weight gradient, in training); the fused training encoder keeps only a2 = ca[1] in row-major form
            ws.cx.append(e(rows, c.kp) if (c.j < n - 1 and (c.j > 0 or train) and not fused_t) else None)
            ws.cz.append(e(rows, c.cout) if (train and c.j < n - 1 and not fused_t) else None)
            ws.ca.append(e(rows, c.cout) if not (fused_t and c.j != 1) else None)
        ws.x0 = e(M, self.H)

    def alloc_decode(self, B: int) -> SimpleNamespace:
        dev, H = self.dev, self.H
        ws = SimpleNamespace(B=B)
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)
        self._alloc_embed(ws, B, train=False)
        ws.xa, ws.xb, ws.xn = e(B, H), e(B, H), e(B, H)
        ws.q, ws.ao, ws.h = e(B, self.QD), e(B, self.QD), e(B, self.F)
        ws.xf, ws.pv = e(B, H), e(B, self.Vh)
        ws.logits = torch.empty(B, self.A, dtype=F32, device=dev)
        ws.vlogits = torch.empty(B, self.NL, dtype=F32, device=dev)
        S = self.cfg.max_seq_length
        ws.kcache = [torch.zeros(B, S, self.Nkv, self.D, dtype=BF16, device=dev) for _ in range(self.L)]
        ws.vcache = [torch.zeros(B, S, self.Nkv, self.D, dtype=BF16, device=dev) for _ in range(self.L)]
        return ws

    def reset_carry(self, ws):
        """initialize_carry (attention.py:102-106): zero KV ring buffers."""
        for t in ws.kcache + ws.vcache:
            t.zero_()

    def alloc_train(self, Bm: int, T: int) -> SimpleNamespace:
        dev, H, F = self.dev, self.H, self.F
        M = Bm * T
        ws = SimpleNamespace(Bm=Bm, T=T, M=M)
        e = lambda *s: torch.empty(*s, dtype=BF16, device=dev)
        f = lambda *s: torch.empty(*s, dtype=F32, device=dev)
        self._alloc_embed(ws, M, train=True)
        L = self.L
        ws.x = [ws.x0] + [e(M, H) for _ in range(L)]
        ws.xn1 = [e(M, H) for _ in range(L)]
        ws.qkv = [e(M, self.QKV) for _ in range(L)]
        ws.o = [e(M, self.QD) for _ in range(L)]
        ws.lse = [f(M, self.Nq) for _ in range(L)]
        ws.x1 = [e(M, H) for _ in range(L)]
        ws.xn2 = [e(M, H) for _ in range(L)]
        ws.h = [e(M, F) for _ in range(L)]
        if not self.fused_glu_train:
            ws.u = [e(M, F) for _ in range(L)]
            ws.g = [e(M, F) for _ in range(L)]
        ws.xf, ws.pv = e(M, H), e(M, self.Vh)
        if self.fused_value_train:
            ws.zv, ws.vval = None, f(M)
            ws.vb_ws = torch.empty(_lib.lib.mpx_value_bwd_fused_workspace_bytes(self.Vh), dtype=torch.uint8, device=dev)
        else:
            ws.zv = e(M, self.Vh)
        ws.logits, ws.vlogits = f(M, self.A), f(M, self.NL)
        ws.pos = torch.arange(T, dtype=I32, device=dev)
        # backward scratch
        ws.dlogits = f(M, self.A)
        ws.dvl = e(M, 64)
        ws.dpv = e(M, self.Vh)
        ws.dxa, ws.dxb, ws.dxn = e(M, H), e(M, H), e(M, H)
        ws.dh, ws.dug = (None if self.fused_glu_train else e(M, F)), e(M, 2 * F)
        ws.dqkv, ws.do = e(M, self.QKV), e(M, self.QD)
        ws.delta, ws.dq_acc = f(M, self.Nq), f(M, self.QD)
        if self.fused_obs_train:                       # only dz1 (operand of the one-hot weight gradient) is row-major
            ws.dca = [e(M * c.oh * c.ow, c.cout) if c.j == 0 else None for c in self.conv[:-1]]
            ws.dcx = [None for c in self.conv[:-1]]
        else:
            ws.dca = [e(M * c.oh * c.ow, c.cout) for c in self.conv[:-1]]
            ws.dcx = [e(M * c.oh * c.ow, c.kp) if c.j > 0 else None for c in self.conv[:-1]]
        ws.losses = f(4)          # value_loss, actor_loss, entropy_loss, (spare)
        ws.loss_ws = torch.empty(max(_lib.lib.mpx_hlgauss_loss_workspace_bytes(M),
                                     _lib.lib.mpx_ppo_policy_loss_workspace_bytes(M)), dtype=torch.uint8, device=dev)
        return ws

    # ------------------------------------------------------------------------------------------ embed (network.py:303-309)
    def _embed_fwd(self, ws, M: int, obs, reward, last_action, task_ids, out, train: bool):
        p, w = self.p, self.w
        cfg = self.cfg
        n = len(self.conv)
        if self.flat:
            ops.flat_embed_fwd(obs, p["obs_encoder.embedding.kernel"], p["obs_encoder.embedding.bias"], ws.fe, M, self.cells)
            ops.linear(ws.fe, w["wt_fd"], bias=w["b_fd"], out=ws.oe)
            t_table = p.views.get("task_embedder.embedding_table") if task_ids is not None else None
            ops.embed_combine(ws.oe, reward, last_action, task_ids if t_table is not None else None,
                              p["reward_encoder.kernel"], p["reward_encoder.bias"], p["action_embedder.embedding_table"],
                              t_table, out=out)
            return out
        if not train and self.fused_obs:          # rollout: encoder + embedding sum in one kernel, nothing saved
            t_table = p.views.get("task_embedder.embedding_table") if task_ids is not None else None
            c0 = self.conv[0]
            return ops.obs_embed_fused(obs, cfg.obs_max_value, self.obs_image, reward, last_action,
                                       task_ids if t_table is not None else None, p["reward_encoder.kernel"],
                                       p["reward_encoder.bias"], p["action_embedder.embedding_table"], t_table, out, M,
                                       c0.ih, c0.iw)
        if train and self.fused_obs_train:        # training: one kernel, z1 / a1 / z2 / a2 saved for the backward
            t_table = p.views.get("task_embedder.embedding_table") if task_ids is not None else None
            c0, t = self.conv[0], ws.obs_t
            return ops.obs_train_fwd(obs, cfg.obs_max_value, self.obs_image, reward, last_action,
                                     task_ids if t_table is not None else None, p["reward_encoder.kernel"],
                                     p["reward_encoder.bias"], p["action_embedder.embedding_table"], t_table,
                                     t["z1t"], t["a1t"], t["z2t"], ws.ca[1], out, M, c0.ih, c0.iw)
        prev = None
        for c in self.conv:
            last = c.j == n - 1
            if c.j == 0 and not last:
                ops.conv1_onehot_fwd(obs, w["w_c0"], w["b_c0"], ws.ca[0], M, c.ih, c.iw, cfg.obs_max_value, c.kh, c.kw,
                                     c.sh, c.sw, c.cout, z_out=ws.cz[0] if train else None, gelu=True)
                prev = ws.ca[0]
                continue
            elif c.j == 0:
                raise _lib.MpxError("a single-layer grid_cnn encoder is not supported")
            elif last:
                xin = prev.view(M, c.kp) if c.kp == c.k else None
                if xin is None:
                    raise _lib.MpxError("last conv layer needs kh*kw*cin % 8 == 0")
            else:
                ops.im2col(prev, ws.cx[c.j], M, c.ih, c.iw, c.cin, c.kh, c.kw, c.sh, c.sw)
                xin = ws.cx[c.j]
            rows = M * c.oh * c.ow
            ops.linear(xin, w[f"wt_c{c.j}"], bias=w[f"b_c{c.j}"], act=0 if last else 1, out=ws.ca[c.j],
                       y_pre=ws.cz[c.j] if (train and not last) else None, M=rows)
            prev = ws.ca[c.j]
        t_table = p.views.get("task_embedder.embedding_table") if task_ids is not None else None
        ops.embed_combine(prev, reward, last_action, task_ids if t_table is not None else None,
                          p["reward_encoder.kernel"], p["reward_encoder.bias"], p["action_embedder.embedding_table"],
                          t_table, out=out)
        return out

    def _embed_bwd(self, ws, M: int, dx0, reward, last_action, task_ids, obs, matrices_ready=None):
        p, w = self.p, self.w
        gt = p.gviews.get("task_embedder.embedding_table") if task_ids is not None else None
        ops.embed_bwd(dx0, reward, last_action, task_ids if gt is not None else None, self.A, self.cfg.task_count,
                      p.g("reward_encoder.kernel"), p.g("reward_encoder.bias"),
                      p.g("action_embedder.embedding_table"), gt)
        # Every 2-D leaf outside obs_encoder / reward_encoder / value_head (the Muon partition, optimizer.py:42-69) now
        # has its complete gradient; only the observation encoder's backward is left.  The caller's hook (gradient
        # all-reduce of that partition + the Newton-Schulz phase of the optimizer) runs on the side stream under it.
        self.matrices_event = self._side_run(matrices_ready, self.side2) if matrices_ready is not None else None
        if self.flat:
            ops.colsum(dx0, p.g("obs_encoder.dense.bias"), M=M, N=self.H, ld=self.H)
            ops.linear_wgrad(ws.fe, dx0, p.g("obs_encoder.dense.kernel"), M=M, K=self.fk, N=self.H, ldx=self.fkp,
                             ldy=self.H, lddw=self.H)
            ops.linear(dx0, w["w_fd"], out=ws.dfe)
            ops.flat_embed_bwd(obs, ws.dfe, p.g("obs_encoder.embedding.kernel"), p.g("obs_encoder.embedding.bias"), M,
                               self.cells)
            return
        if self.fused_obs_train:
            # conv3: bias / weight gradient from dx0 and the saved a2; one kernel for both data gradients (+ the conv1 /
            # conv2 bias gradients); tap-wise conv2 weight gradient; one-hot conv1 weight gradient.  No im2col / col2im.
            c0, c1, c2, t = self.conv[0], self.conv[1], self.conv[2], ws.obs_t
            ops.colsum(dx0, p.g("obs_encoder.layers.2.bias"), M=M, N=c2.cout, ld=c2.cout)
            ops.linear_wgrad(ws.ca[1].view(M, c2.kp), dx0, p.g("obs_encoder.layers.2.kernel").view(c2.k, c2.cout), M=M,
                             K=c2.k, N=c2.cout, ldx=c2.kp, ldy=c2.cout, lddw=c2.cout)
            ops.obs_train_bwd(dx0, t["z1t"], t["z2t"], self.obs_image, t["dz2t"], ws.dca[0],
                              p.g("obs_encoder.layers.0.bias"), p.g("obs_encoder.layers.1.bias"), M, t["ws"])
            ops.obs_conv2_wgrad(t["a1t"], t["dz2t"], p.g("obs_encoder.layers.1.kernel"), M, t["ws"])
            ops.conv1_onehot_wgrad(obs, ws.dca[0], p.g("obs_encoder.layers.0.kernel").view(c0.k, c0.cout), M, c0.ih, c0.iw,
                                   self.cfg.obs_max_value, c0.kh, c0.kw, c0.sh, c0.sw, c0.cout)
            return
        # Data-gradient chain (GEMM -> col2im -> GEMM ...) and each layer's bias / weight gradients, all on this stream:
        # both are HBM-bound, so running the weight gradients on the side stream gained nothing (round 1, trip 79).
        n = len(self.conv)
        dy = dx0                      # cotangent of the last conv's output (M, H)
        fused_gelu = False
        for c in reversed(self.conv):
            last = c.j == n - 1
            rows = M * c.oh * c.ow
            if not last and not fused_gelu:
                ops.gelu_bwd(dy, ws.cz[c.j], out=dy)          # dz in place
            fused_gelu = False

            def _wg(c=c, dy=dy, last=last, rows=rows):
                ops.colsum(dy, p.g(f"obs_encoder.layers.{c.j}.bias"), M=rows, N=c.cout, ld=c.cout)
                if c.j == 0 and self.onehot_wgrad:     # one-hot operand tiles are generated inside the wgrad kernel
                    ops.conv1_onehot_wgrad(obs, dy, p.g("obs_encoder.layers.0.kernel").view(c.k, c.cout), M, c.ih, c.iw,
                                           self.cfg.obs_max_value, c.kh, c.kw, c.sh, c.sw, c.cout)
                    return
                if c.j == 0:      # one-hot im2col rows are only needed here, for the weight gradient
                    ops.obs_onehot_im2col(obs, ws.cx[0], M, c.ih, c.iw, self.cfg.obs_max_value, c.kh, c.kw, c.sh, c.sw)
                xin = ws.ca[c.j - 1].view(M, c.kp) if last else ws.cx[c.j]
                ops.linear_wgrad(xin, dy, p.g(f"obs_encoder.layers.{c.j}.kernel").view(c.k, c.cout), M=rows, K=c.k,
                                 N=c.cout, ldx=c.kp, ldy=c.cout, lddw=c.cout)
            _wg()        # on this stream: both sides of the encoder backward are HBM-bound, overlapping them gained nothing
            if c.j == 0:
                break
            if last:
                ops.linear(dy, w[f"w_c{c.j}"], out=ws.dca[c.j - 1].view(M, c.k), M=rows)
            else:
                ops.linear(dy, w[f"w_c{c.j}"], out=ws.dcx[c.j], M=rows)
                # the layer below is never the last one: its gelu transpose rides along with the scatter-sum
                ops.col2im(ws.dcx[c.j], ws.dca[c.j - 1], M, c.ih, c.iw, c.cin, c.kh, c.kw, c.sh, c.sw, z=ws.cz[c.j - 1])
                fused_gelu = True
            dy = ws.dca[c.j - 1]

    # ------------------------------------------------------------------------------------------ decode step
    def decode_step(self, ws, obs, reward, last_action, pos, action_mask=None, task_ids=None, value_out=None,
                    sample=None, hidden_out=None):
        """One rollout step for B agents: TransformerActorCritic.__call__ with carry (network.py:300-343) on
        add_seq_dim(timestep).  pos is the (B,) int32 device array ts.time.  Leaves action logits in ws.logits
        and the HL-Gauss logits in ws.vlogits; the KV caches in ws are updated in place.  value_out (B,) f32, when
        given, receives HlGaussValue.get_value of those logits (values.py:43-45; fused with the value path).
        sample = dict(action=(B,) i32, log_prob=(B,) f32, seed=, stream_id=, stream_offset=, [gumbel_in=],
        [want_logits=True]) draws the action inside the policy-head kernel (sample_actions, network.py:345-351).
        hidden_out (B,H) bf16: the trunk output x is left there and the value path is NOT run - the caller evaluates
        it later for many steps at once with values_from_hidden (the value is not an input of the next step)."""
        p, w, B = self.p, self.w, ws.B
        S = self.cfg.max_seq_length
        x = self._embed_fwd(ws, B, obs, reward, last_action, task_ids, ws.xa, train=False)
        other = ws.xb
        for i in range(self.L):
            pre = f"layers.{i}."
            if self.fused_attn_decode:                        # pre-norm, projection, RoPE, ring write and attention: one kernel
                # Only the attention / GLU kernels launch programmatically (default mask): the encoder and head kernels
                # of every step are full stream barriers, so the ring slots of earlier steps are complete and visible
                # before this launch can begin - their stream may start under the previous kernel's tail.
                ops.attn_decode_fused(x, p[pre + "history_norm.scale"], w[pre + "wt_qkv"], pos, self.timescale,
                                      ws.kcache[i], ws.vcache[i], self.Nq, self.Nkv, self.D, out=ws.ao)
            elif self.H == 128:                               # pre-norm fused into the q/k/v GEMM (shared-memory prologue)
                ops.qkv_rope(x, w[pre + "wt_qkv"], pos, self.timescale, self.Nq, self.Nkv, self.D, q=ws.q,
                             k=ws.kcache[i], v=ws.vcache[i], k_row_stride=S * self.KD, v_row_stride=S * self.KD,
                             cache_S=S, norm_scale=p[pre + "history_norm.scale"])
            else:
                ops.rmsnorm_fwd(x, p[pre + "history_norm.scale"], out=ws.xn)
                ops.qkv_rope(ws.xn, w[pre + "wt_qkv"], pos, self.timescale, self.Nq, self.Nkv, self.D, q=ws.q,
                             k=ws.kcache[i], v=ws.vcache[i], k_row_stride=S * self.KD, v_row_stride=S * self.KD,
                             cache_S=S)
            if not self.fused_attn_decode:
                ops.attn_decode(ws.q, ws.kcache[i], ws.vcache[i], pos, self.Nq, self.Nkv, self.D, out=ws.ao)
            if hidden_out is not None and i == self.L - 1:
                other = hidden_out                            # the block's output goes straight to the caller's buffer
            if self.fused_glu and self.QD == 128:
                # out projection + residual, ffn_norm, GLU and its residual in ONE kernel (x1 stays in shared memory)
                if hidden_out is not None and sample is not None and i == self.L - 1:
                    # ... and, on the last layer of a rollout step, the policy head on the finished rows (one launch less)
                    ops.glu_fused_head(ws.ao, w[pre + "wt_ug64"], w[pre + "wt_d"], x, other, self.F,
                                       p[pre + "ffn_norm.scale"], w[pre + "wt_o"], p["output_norm.scale"],
                                       p["action_embedder.embedding_table"], sample["action"], sample["log_prob"],
                                       seed=sample.get("seed", 0), stream_id=sample.get("stream_id", 0),
                                       stream_offset=sample.get("stream_offset"), mask=action_mask,
                                       gumbel_in=sample.get("gumbel_in"),
                                       logits_out=ws.logits if sample.get("want_logits", True) else None)
                    return ws.logits, None
                ops.glu_fused(ws.ao, w[pre + "wt_ug64"], w[pre + "wt_d"], x, other, self.F,
                              norm_scale=p[pre + "ffn_norm.scale"], wt_o=w[pre + "wt_o"])
                x, other = other, x
                continue
            ops.linear(ws.ao, w[pre + "wt_o"], residual=x, out=other)
            x, other = other, x
            if self.fused_glu:                                # ffn_norm is applied inside the kernel
                ops.glu_fused(x, w[pre + "wt_ug64"], w[pre + "wt_d"], x, other, self.F,
                              norm_scale=p[pre + "ffn_norm.scale"])
            else:
                ops.rmsnorm_fwd(x, p[pre + "ffn_norm.scale"], out=ws.xn)
                ops.glu_up(ws.xn, w[pre + "wt_ug"], self.F, h=ws.h)
                ops.linear(ws.h, w[pre + "wt_d"], residual=x, out=other)
            x, other = other, x
        if hidden_out is not None:
            if sample is not None:
                ops.policy_head_sample(x, p["action_embedder.embedding_table"], sample["action"], sample["log_prob"],
                                       seed=sample.get("seed", 0), stream_id=sample.get("stream_id", 0),
                                       stream_offset=sample.get("stream_offset"), mask=action_mask,
                                       gumbel_in=sample.get("gumbel_in"), norm_scale=p["output_norm.scale"],
                                       logits_out=ws.logits if sample.get("want_logits", True) else None)
            return ws.logits, None
        if value_out is not None and self.fused_value:
            # both head kernels apply output_norm themselves: no normalised copy of x is materialised
            if sample is not None:
                ops.policy_head_sample(x, p["action_embedder.embedding_table"], sample["action"], sample["log_prob"],
                                       seed=sample.get("seed", 0), stream_id=sample.get("stream_id", 0),
                                       stream_offset=sample.get("stream_offset"), mask=action_mask,
                                       gumbel_in=sample.get("gumbel_in"), norm_scale=p["output_norm.scale"],
                                       logits_out=ws.logits if sample.get("want_logits", True) else None)
            else:
                ops.rmsnorm_fwd(x, p["output_norm.scale"], out=ws.xf)
                ops.policy_logits(ws.xf, p["action_embedder.embedding_table"], action_mask, out=ws.logits)
            ops.value_head_fused(x, w["wt_vm"], w["b_vm"], w["wt_vh_hilo"], p["value_head.dense.bias"], self.centers,
                                 value_out, vlogits=ws.vlogits, norm_scale=p["output_norm.scale"])
            return ws.logits, ws.vlogits
        self._heads_fwd(x, action_mask, ws.xf, ws.logits, ws.pv, None, ws.vlogits)
        if value_out is not None:
            ops.hlgauss_value(ws.vlogits, self.centers, out=value_out)
        if sample is not None:
            gum = sample.get("gumbel_in")
            if gum is None:
                gum = ops.gumbel_noise(torch.empty_like(ws.logits), sample.get("seed", 0), sample.get("stream_id", 0),
                                       sample.get("stream_offset"))
            ops.policy_sample(ws.logits, gum, action=sample["action"], log_prob=sample["log_prob"])
        return ws.logits, ws.vlogits

    def values_from_hidden(self, hidden, value_out):
        """HlGaussValue.get_value(value_head(gelu(value_mlp(output_norm(x))))) (network.py:321,335-341, values.py:40-45)
        for any number of trunk outputs at once: hidden (M,H) bf16 -> value_out (M,) f32.  Requires the fused kernel."""
        if not self.fused_value:
            raise _lib.MpxError("values_from_hidden needs the fused value path (H == 128)")
        p, w = self.p, self.w
        ops.value_head_fused(hidden, w["wt_vm"], w["b_vm"], w["wt_vh_hilo"], p["value_head.dense.bias"], self.centers,
                             value_out, norm_scale=p["output_norm.scale"])

    def _heads_fwd(self, x, action_mask, xf, logits, pv, zv, vlogits, vval=None):
        """network.py:321-341.  vval (M,) f32: take the one-kernel value path (training; nothing kept for the backward)."""
        p, w = self.p, self.w
        ops.rmsnorm_fwd(x, p["output_norm.scale"], out=xf)
        ops.policy_logits(xf, p["action_embedder.embedding_table"], action_mask, out=logits)
        if vval is not None:                           # value MLP + fp32 head in one kernel
            ops.value_head_fused(xf, w["wt_vm"], w["b_vm"], w["wt_vh_hilo"], p["value_head.dense.bias"], self.centers,
                                 vval, vlogits=vlogits)
            return
        ops.linear(xf, w["wt_vm"], bias=w["b_vm"], act=1, out=pv, y_pre=zv)
        ops.linear_f32w(pv, w["wt_vh_hilo"], p["value_head.dense.bias"], vlogits, self.NL)

    # ------------------------------------------------------------------------------------------ training forward
    def forward_train(self, ws, obs, last_rewards, last_actions, action_mask=None, task_ids=None):
        """Full-window pass (ppo_loss's model call, train.py:180-192): obs (Bm,T,...) u8, positions arange(T)."""
        p, w, M = self.p, self.w, ws.M
        x = self._embed_fwd(ws, M, obs, last_rewards, last_actions, task_ids, ws.x[0], train=True)
        QD, KD, QKV = self.QD, self.KD, self.QKV
        for i in range(self.L):
            pre = f"layers.{i}."
            ops.rmsnorm_fwd(ws.x[i], p[pre + "history_norm.scale"], out=ws.xn1[i])
            qkv = ws.qkv[i]
            ops.qkv_rope(ws.xn1[i], w[pre + "wt_qkv"], ws.pos, self.timescale, self.Nq, self.Nkv, self.D,
                         q=qkv, k=qkv[:, QD:], v=qkv[:, QD + KD:], q_row_stride=QKV, k_row_stride=QKV,
                         v_row_stride=QKV, pos_len=ws.T)
            ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], ws.Bm, ws.T, self.Nq, self.Nkv, self.D, QKV, QKV,
                               QKV, o=ws.o[i], lse=ws.lse[i])
            ops.linear(ws.o[i], w[pre + "wt_o"], residual=ws.x[i], out=ws.x1[i])
            ops.rmsnorm_fwd(ws.x1[i], p[pre + "ffn_norm.scale"], out=ws.xn2[i])
            if self.fused_glu_train:
                ops.glu_fused(ws.xn2[i], w[pre + "wt_ug64"], w[pre + "wt_d"], ws.x1[i], ws.x[i + 1], self.F, h_out=ws.h[i])
            else:
                ops.glu_up(ws.xn2[i], w[pre + "wt_ug"], self.F, save=True, h=ws.h[i], u=ws.u[i], g=ws.g[i])
                ops.linear(ws.h[i], w[pre + "wt_d"], residual=ws.x1[i], out=ws.x[i + 1])
        if self.fused_value_train:
            self._heads_fwd(ws.x[self.L], action_mask, ws.xf, ws.logits, None, None, ws.vlogits, vval=ws.vval)
        else:
            self._heads_fwd(ws.x[self.L], action_mask, ws.xf, ws.logits, ws.pv, ws.zv, ws.vlogits)
        return ws.vlogits, ws.logits

    # ------------------------------------------------------------------------------------------ side stream
    # Weight-gradient GEMMs only feed the optimizer, so they run on a second stream next to the data-gradient chain
    # (HBM-bound wgrads overlap the compute-bound attention backward and the epilogue-bound dgrad GEMMs).  They share
    # one split-K workspace, which the side stream's program order protects.  Every buffer a pending wgrad reads is
    # guarded by an event the main stream waits on before the next kernel that overwrites it.
    def _side_run(self, fn, stream=None):
        """Enqueue fn on the side stream (or `stream`), ordered after everything already on the current stream;
        returns an event recorded on that stream after fn."""
        side = self.side if stream is None else stream
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            fn()
            done = torch.cuda.Event()
            done.record(side)
        return done

    @staticmethod
    def _wait(ev):
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    # ------------------------------------------------------------------------------------------ loss + backward
    def loss_and_backward(self, ws, batch: dict, hyp, progress: float, entropy_coef_dev=None, matrices_ready=None):
        """ppo_loss (train.py:162-222) on the activations left by forward_train, and its full transpose.
        Gradients are ACCUMULATED into self.p.grad (zero it first).  ws.losses = (value, actor, entropy_loss).
        matrices_ready: optional callable enqueued on the side stream as soon as the gradients of the Muon-labelled
        matrices (p.grad[p.n_adam:]) are final; self.matrices_event then holds its completion event (wait on it)."""
        p, w, M = self.p, self.w, ws.M
        cfg = self.cfg
        H, F, QD, KD, QKV, NL, Vh = self.H, self.F, self.QD, self.KD, self.QKV, self.NL, self.Vh
        ent_coef = hyp.entropy_coef + progress * (hyp.entropy_coef_end - hyp.entropy_coef)
        # ---- losses and their logits gradients
        ops.hlgauss_loss(ws.vlogits, batch["targets"], self.supports, cfg.value_min, cfg.value_max, cfg.value_sigma,
                         grad_scale=hyp.vf_coef, loss=ws.losses[0:1], dlogits_bf16=ws.dvl, ws=ws.loss_ws)
        ops.ppo_policy_loss(ws.logits, batch["actions"], batch["log_prob"], batch["advantages"], hyp.vf_clip, ent_coef,
                            losses=ws.losses[1:3], dlogits=ws.dlogits, ws=ws.loss_ws,
                            entropy_coef_dev=entropy_coef_dev)
        # ---- value head (values.py:34,40-41) and value MLP (network.py:335-338)
        def _wg_vhead():                          # dvl / dpv are not rewritten in this pass
            ops.linear_wgrad(ws.pv, ws.dvl, p.g("value_head.dense.kernel"), M=M, K=Vh, N=NL, ldx=Vh, ldy=64, lddw=NL)
            ops.colsum(ws.dvl, p.g("value_head.dense.bias"), M=M, N=NL, ld=64)
        if self.fused_value_train:
            # one kernel: recompute the pre-activation, pv / dz for the weight gradients, d xf, value-MLP bias gradient
            ops.value_bwd_fused(ws.xf, ws.dvl, w["wt_vm"], w["b_vm"], w["w_vh_pad"], ws.pv, ws.dpv, ws.dxn,
                                p.g("value_mlp.bias"), ws=ws.vb_ws)
            self._side_run(_wg_vhead)
            self._side_run(lambda: ops.linear_wgrad(ws.xf, ws.dpv, p.g("value_mlp.kernel"), M=M, K=H, N=Vh, ldx=H, ldy=Vh,
                                                    lddw=Vh))
        else:
            self._side_run(_wg_vhead)
            ops.linear(ws.dvl, w["w_vh_pad"], out=ws.dpv)
            ops.gelu_bwd(ws.dpv, ws.zv, out=ws.dpv)

            def _wg_vmlp():
                ops.linear_wgrad(ws.xf, ws.dpv, p.g("value_mlp.kernel"), M=M, K=H, N=Vh, ldx=H, ldy=Vh, lddw=Vh)
                ops.colsum(ws.dpv, p.g("value_mlp.bias"), M=M, N=Vh, ld=Vh)
            self._side_run(_wg_vmlp)
            ops.linear(ws.dpv, w["w_vm"], out=ws.dxn)                  # d xf from the value path
        # ---- tied policy head (network.py:323) and output norm (:321)
        ops.policy_head_bwd(ws.dlogits, ws.xf, p["action_embedder.embedding_table"],
                            p.g("action_embedder.embedding_table"), dx_other=ws.dxn, out=ws.dxa)
        dx, spare = ws.dxb, ws.dxa
        ops.rmsnorm_bwd(ws.dxa, ws.x[self.L], p["output_norm.scale"], dscale=p.g("output_norm.scale"), out=dx)
        # ---- transformer blocks in reverse (network.py:158-172)
        ev_updown = ev_out = ev_qkv = None          # side-stream completion of the previous layer's wgrads
        for i in reversed(range(self.L)):
            pre = f"layers.{i}."
            dx_in = dx
            if not self.fused_glu_train:
                ops.linear(dx, w[pre + "w_d"], out=ws.dh)
            ev_down = self._side_run(lambda: ops.linear_wgrad(ws.h[i], dx_in, p.g(pre + "ffn.down_proj.kernel"), M=M, K=F,
                                                              N=H, ldx=F, ldy=H, lddw=H))
            self._wait(ev_updown)                    # dug is still being read by the previous layer's up/gate wgrads
            if self.fused_glu_train:
                ops.glu_bwd_fused(ws.xn2[i], dx, w[pre + "wt_ug64"], w[pre + "wt_d"], self.F, dug=ws.dug, dxn=ws.dxn)
            else:
                ops.glu_bwd(ws.dh, ws.u[i], ws.g[i], out=ws.dug)

            def _wg_updown(i=i, pre=pre):
                ops.linear_wgrad(ws.xn2[i], ws.dug, p.g(pre + "ffn.up_proj.kernel"), M=M, K=H, N=F, ldx=H, ldy=2 * F,
                                 lddw=F)
                ops.linear_wgrad(ws.xn2[i], ws.dug[:, F:], p.g(pre + "ffn.up_gate.kernel"), M=M, K=H, N=F, ldx=H,
                                 ldy=2 * F, lddw=F)
            ev_updown = self._side_run(_wg_updown)
            if not self.fused_glu_train:
                ops.linear(ws.dug, w[pre + "w_ug"], out=ws.dxn)
            self._wait(ev_out)                       # `spare` was the d x1 the previous layer's out-proj wgrad reads
            ops.rmsnorm_bwd(ws.dxn, ws.x1[i], p[pre + "ffn_norm.scale"], dscale=p.g(pre + "ffn_norm.scale"), dres=dx,
                            out=spare)
            dx, spare = spare, dx                                       # dx = d x1
            dx_mid = dx
            ops.linear(dx, w[pre + "w_o"], out=ws.do)
            ev_out = self._side_run(lambda: ops.linear_wgrad(ws.o[i], dx_mid, p.g(pre + "history.out.kernel").view(QD, H),
                                                             M=M, K=QD, N=H, ldx=QD, ldy=H, lddw=H))
            qkv, dqkv = ws.qkv[i], ws.dqkv
            self._wait(ev_qkv)                       # dqkv is still being read by the previous layer's q/k/v wgrads
            ops.attn_train_bwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], ws.o[i], ws.do, ws.lse[i], ws.pos, self.timescale,
                               ws.Bm, ws.T, self.Nq, self.Nkv, self.D, QKV, QKV, QKV, dqkv, dqkv[:, QD:],
                               dqkv[:, QD + KD:], QKV, QKV, QKV, delta=ws.delta, dq_acc=ws.dq_acc, pos_len=ws.T)

            def _wg_qkv(i=i, pre=pre, dqkv=dqkv):
                ops.linear_wgrad(ws.xn1[i], dqkv, p.g(pre + "history.query_proj.kernel").view(H, QD), M=M, K=H, N=QD,
                                 ldx=H, ldy=QKV, lddw=QD)
                ops.linear_wgrad(ws.xn1[i], dqkv[:, QD:], p.g(pre + "history.key_proj.kernel").view(H, KD), M=M, K=H,
                                 N=KD, ldx=H, ldy=QKV, lddw=KD)
                ops.linear_wgrad(ws.xn1[i], dqkv[:, QD + KD:], p.g(pre + "history.value_proj.kernel").view(H, KD), M=M,
                                 K=H, N=KD, ldx=H, ldy=QKV, lddw=KD)
            ev_qkv = self._side_run(_wg_qkv)
            ops.linear(dqkv, w[pre + "w_qkv"], out=ws.dxn)
            self._wait(ev_down)                      # `spare` is the d x[i+1] this layer's down-proj wgrad reads
            ops.rmsnorm_bwd(ws.dxn, ws.x[i], p[pre + "history_norm.scale"], dscale=p.g(pre + "history_norm.scale"),
                            dres=dx, out=spare)
            dx, spare = spare, dx                                       # dx = d x[i]
        # the embedding backward uses the same split-K workspace and overwrites `spare`: join the side stream first
        self._wait(ev_qkv)
        self._embed_bwd(ws, M, dx, batch["last_rewards"], batch["last_actions"], batch.get("task_ids"), batch["obs"],
                        matrices_ready=matrices_ready)
        return ws.losses
