"""Parameter store: one flat fp32 buffer (+ a flat gradient buffer) with named views that follow the reference's
nnx pytree paths, so weights can be handed to / from the JAX side by name.

Initialisers restate network.py:39-52,190-193 (glorot-uniform layer kernels, N(0,1)*0.01 embedding tables,
flax defaults - lecun-normal kernels, zero biases - elsewhere, ones for norm scales).
"""

from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .config import ModelConfig


def conv_out(n: int, k: int, s: int) -> int:
    return (n - k) // s + 1


def param_specs(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    H, D, Nq, Nkv, F = cfg.hidden_features, cfg.head_dim, cfg.num_heads, cfg.num_kv_heads, cfg.ffn_size
    specs: List[Tuple[str, Tuple[int, ...], str]] = []
    specs.append(("reward_encoder.kernel", (1, H), "lecun"))
    specs.append(("reward_encoder.bias", (H,), "zeros"))
    specs.append(("action_embedder.embedding_table", (cfg.action_dim, H), "embed"))
    C = sum(cfg.obs_max_value)
    if cfg.obs_type == "grid_cnn":
        cin = C
        chans = list(cfg.cnn_channels) + [H]
        for j, ((kh, kw), ch) in enumerate(zip(cfg.cnn_kernels, chans)):
            specs.append((f"obs_encoder.layers.{j}.kernel", (kh, kw, cin, ch), "lecun"))
            specs.append((f"obs_encoder.layers.{j}.bias", (ch,), "zeros"))
            cin = ch
    elif cfg.obs_type == "grid_flattened":
        specs.append(("obs_encoder.embedding.kernel", (C, 4), "lecun"))
        specs.append(("obs_encoder.embedding.bias", (4,), "zeros"))
        specs.append(("obs_encoder.dense.kernel", (4 * cfg.obs_shape[0] * cfg.obs_shape[1], H), "lecun"))
        specs.append(("obs_encoder.dense.bias", (H,), "zeros"))
    else:
        raise ValueError(f"unsupported obs_type {cfg.obs_type}")
    if cfg.task_count > 1:
        specs.append(("task_embedder.embedding_table", (cfg.task_count, H), "embed"))
    for i in range(cfg.num_layers):
        pre = f"layers.{i}."
        specs.append((pre + "history_norm.scale", (H,), "ones"))
        specs.append((pre + "history.key_proj.kernel", (H, Nkv, D), "glorot"))
        specs.append((pre + "history.value_proj.kernel", (H, Nkv, D), "glorot"))
        specs.append((pre + "history.query_proj.kernel", (H, Nq, D), "glorot"))
        specs.append((pre + "history.out.kernel", (Nq, D, H), "glorot_out"))
        specs.append((pre + "ffn_norm.scale", (H,), "ones"))
        specs.append((pre + "ffn.up_proj.kernel", (H, F), "glorot"))
        specs.append((pre + "ffn.up_gate.kernel", (H, F), "glorot"))
        specs.append((pre + "ffn.down_proj.kernel", (F, H), "glorot"))
    specs.append(("output_norm.scale", (H,), "ones"))
    vin = H
    if cfg.value_hidden_dim is not None:
        specs.append(("value_mlp.kernel", (H, cfg.value_hidden_dim), "lecun"))
        specs.append(("value_mlp.bias", (cfg.value_hidden_dim,), "zeros"))
        vin = cfg.value_hidden_dim
    specs.append(("value_head.dense.kernel", (vin, cfg.value_n_logits), "lecun"))
    specs.append(("value_head.dense.bias", (cfg.value_n_logits,), "zeros"))
    return specs


MUON_EXCLUDE_TOP = {"value_head", "action_encoder", "reward_encoder", "obs_encoder", "obs_decoder"}   # optimizer.py:42-48


def is_muon_param(name: str, shape) -> bool:
    """muon_dim_numbers_fn (optimizer.py:51-69): 2-D leaves whose top-level module is not excluded."""
    return len(shape) == 2 and name.split(".")[0] not in MUON_EXCLUDE_TOP


def _init(shape, kind: str, gen: torch.Generator) -> torch.Tensor:
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "ones":
        return torch.ones(shape)
    if kind == "embed":
        return torch.randn(shape, generator=gen) * 0.01
    n = 1
    for s in shape:
        n *= s
    if kind == "glorot":          # (in, *out)
        fan_in, fan_out = shape[0], n // shape[0]
    elif kind == "glorot_out":    # (*in, out)
        fan_in, fan_out = n // shape[-1], shape[-1]
    else:                         # lecun: (*receptive, in, out)
        fan_in, fan_out = n // shape[-1], shape[-1]
    if kind.startswith("glorot"):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=gen) * 2 - 1) * lim
    return torch.randn(shape, generator=gen) * math.sqrt(1.0 / fan_in)


def assign_muon_owners(sizes: List[int], world: int) -> List[int]:
    """Owner rank of every Muon matrix: longest-processing-time greedy (largest matrix first to the least loaded rank;
    ties to the lowest rank), so the Newton-Schulz work and the shard sizes are balanced."""
    load = [0] * world
    owner = [0] * len(sizes)
    for i in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += sizes[i]
    return owner


class ParamStore:
    ALIGN = 4   # floats (16 bytes): kernels read fp32 params with float4 loads

    def __init__(self, cfg: ModelConfig, device, seed: int = 0, bias_noise: float = 0.0, world: int = 1):
        """world > 1 lays the Muon-labelled matrices out in `world` equal-sized shards, one per data-parallel rank (the
        matrices a rank owns are contiguous, the shard is zero-padded to the common size): the optimizer reduce-scatters
        the gradient of that partition, runs Newton-Schulz on the local shard only and all-gathers the updates
        (optimizer.py:25-37,51-69 sharded ZeRO-1 style).  Parameter VALUES do not depend on `world`."""
        self.cfg = cfg
        self.world = world
        specs = param_specs(cfg)
        adam = [sp for sp in specs if not is_muon_param(sp[0], sp[1])]
        muon = [sp for sp in specs if is_muon_param(sp[0], sp[1])]
        numel = lambda shape: math.prod(shape)
        pad = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        owners = assign_muon_owners([pad(numel(sp[1])) for sp in muon], world)
        self.muon_owner: Dict[str, int] = {sp[0]: o for sp, o in zip(muon, owners)}
        # AdamW-labelled leaves first, then the Muon-labelled matrices grouped by owner rank
        self.specs = adam + [sp for r in range(world) for sp, o in zip(muon, owners) if o == r]
        self.offsets: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape, _ in adam:
            self.offsets[name] = (off, shape)
            off += pad(numel(shape))
        self.n_adam = off
        loads = [sum(pad(numel(sp[1])) for sp, o in zip(muon, owners) if o == r) for r in range(world)]
        self.muon_shard = max(loads) if muon else 0           # floats per rank shard (equal on every rank)
        for r in range(world):
            o = self.n_adam + r * self.muon_shard
            for sp, ow in zip(muon, owners):
                if ow == r:
                    self.offsets[sp[0]] = (o, sp[1])
                    o += pad(numel(sp[1]))
        self.numel = off = self.n_adam + world * self.muon_shard
        gen = torch.Generator().manual_seed(seed)
        host = torch.zeros(off, dtype=torch.float32)
        for name, shape, kind in adam + muon:    # fixed draw order: the values do not depend on the shard layout
            o, _ = self.offsets[name]
            t = _init(shape, kind, gen)
            if kind == "zeros" and bias_noise > 0:
                t = torch.randn(shape, generator=gen) * bias_noise
            host[o:o + t.numel()] = t.reshape(-1)
        self.flat = host.to(device)
        self.grad = torch.zeros_like(self.flat)
        self.views = {n: self._view(self.flat, n) for n, _, _ in self.specs}
        self.gviews = {n: self._view(self.grad, n) for n, _, _ in self.specs}

    def muon_table(self, rank: int = None):
        """-> (device uint8 tensor of MuonMat records, nmat, x_total, a_total, max_r, max_c) for mpx_muon_step.
        rank: only the matrices that rank owns (the sharded optimizer); None = all of them."""
        import struct
        recs, xoff, aoff, max_r, max_c = [], 0, 0, 1, 1
        for name, shape, _ in self.specs:
            if not is_muon_param(name, shape) or (rank is not None and self.muon_owner[name] != rank):
                continue
            off, _ = self.offsets[name]
            rows, cols = shape
            tr = 1 if rows > cols else 0
            r, c = (cols, rows) if tr else (rows, cols)
            recs.append(struct.pack("<q6i2q", off, rows, cols, r, c, tr, 0, xoff, aoff))
            xoff += r * c
            aoff += r * r
            max_r, max_c = max(max_r, r), max(max_c, c)
        raw = b"".join(recs)
        dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.flat.device) if recs else None
        return dev, len(recs), xoff, aoff, max_r, max_c

    def _view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        o, shape = self.offsets[name]
        n = 1
        for s in shape:
            n *= s
        return buf[o:o + n].view(shape)

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]

    def g(self, name: str) -> torch.Tensor:
        return self.gviews[name]

    def oracle_dict(self) -> Dict[str, torch.Tensor]:
        """fp32 CPU copies keyed by the reference pytree path (what the oracle / a JAX checkpoint consumes)."""
        return {n: v.detach().float().cpu().clone() for n, v in self.views.items()}

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return {n: v.detach().float().cpu().clone() for n, v in self.gviews.items()}

    def load_dict(self, d: Dict[str, torch.Tensor], strict: bool = True) -> None:
        """Copy named fp32 arrays (reference pytree paths) into the flat buffer.  strict: the key sets must match
        exactly and every shape must be the reference's (what an orbax StandardRestore would enforce)."""
        if strict:
            missing, extra = sorted(set(self.views) - set(d)), sorted(set(d) - set(self.views))
            if missing or extra:
                raise KeyError(f"parameter names differ: missing {missing[:5]}, unexpected {extra[:5]}")
        for n, v in d.items():
            if n not in self.views:
                continue
            v = torch.as_tensor(v)
            if tuple(v.shape) != tuple(self.views[n].shape):
                raise ValueError(f"{n}: shape {tuple(v.shape)} != {tuple(self.views[n].shape)}")
            self.views[n].copy_(v.to(self.flat.device, torch.float32))

    # ------------------------------------------------------------------ weight hand-off (checkpointer.py:15-34)
    # The reference saves nnx.state(model, nnx.Param) with orbax.  Reading orbax/tensorstore trees needs those
    # libraries, so the hand-off format is a flat .npz written on the reference side by
    # scripts/export_reference_params.py ("param/<nnx path>" float32 arrays, "meta/*" env facts).
    def save_npz(self, path: str, **meta) -> None:
        import numpy as np
        arrays = {"param/" + n: v.detach().float().cpu().numpy() for n, v in self.views.items()}
        arrays.update({"meta/" + k: np.asarray(v) for k, v in meta.items()})
        np.savez(path, **arrays)

    def load_npz(self, path: str, strict: bool = True) -> dict:
        """Load parameters exported by scripts/export_reference_params.py (or save_npz); returns the meta entries."""
        import numpy as np
        with np.load(path, allow_pickle=False) as z:
            params = {k[len("param/"):]: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith("param/")}
            meta = {k[len("meta/"):]: z[k] for k in z.files if k.startswith("meta/")}
        self.load_dict(params, strict=strict)
        return meta
