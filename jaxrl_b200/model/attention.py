"""Mirror of mapox_trainer/model/attention.py (AttentionBlock + KVCache) over the B200 kernels.

Same constructor arguments, the same three methods and the same error for an indivisible d_model.  The third
`attention_impl` value the reference's selector seam (attention.py:40-43) gains is "b200"; it is the only one here.
"""

from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch

from .. import ops
from .positional_embeddings import rope_timescale

BF16 = torch.bfloat16


class KVCache(NamedTuple):
    key: torch.Tensor
    value: torch.Tensor


class AttentionBlock:
    def __init__(self, d_model: int, head_dim: int, num_heads: int, num_kv_heads: int, *, max_seq_length: int,
                 rope_max_wavelength: float = 10_000, use_qk_norm: bool = False, attention_impl: Optional[str] = None,
                 dtype=BF16, param_dtype=torch.float32, kernel_init=None, rngs=0, device="cuda"):
        self.d_model, self.head_dim, self.num_heads, self.num_kv_heads = d_model, head_dim, num_heads, num_kv_heads
        self.use_qk_norm = use_qk_norm
        self.max_seq_length = max_seq_length
        self.rope_max_wavelength = rope_max_wavelength
        self.attention_impl = "b200" if attention_impl is None else attention_impl
        self.dtype, self.param_dtype = dtype, param_dtype
        if self.d_model % self.num_heads != 0:
            raise ValueError(
                f"Memory dimension ({self.d_model}) must be divisible by "
                f"'num_heads' heads ({self.num_heads})."
            )
        if self.attention_impl != "b200":
            raise ValueError(f"attention_impl {self.attention_impl!r} is not available here (only 'b200')")
        if use_qk_norm:
            raise NotImplementedError("use_qk_norm is false in every shipped config; not on the B200 path")
        if dtype != BF16:
            raise TypeError("the B200 path is bf16 compute / fp32 params (every shipped config)")
        gen = torch.Generator().manual_seed(int(rngs) if not isinstance(rngs, torch.Generator) else 0)
        init = kernel_init if kernel_init is not None else (lambda shape: torch.randn(shape, generator=gen) * 0.01)
        dev = torch.device(device)
        self.device = dev
        self.query_proj = init((d_model, num_heads, head_dim)).to(dev)
        self.key_proj = init((d_model, num_kv_heads, head_dim)).to(dev)
        self.value_proj = init((d_model, num_kv_heads, head_dim)).to(dev)
        self.out = init((num_heads, head_dim, d_model)).to(dev)
        self._timescale = rope_timescale(head_dim, rope_max_wavelength, dev)
        self.refresh()

    def refresh(self):
        """Re-stage the bf16 operand copies after the fp32 kernels changed."""
        QD, KD = self.num_heads * self.head_dim, self.num_kv_heads * self.head_dim
        wq = self.query_proj.reshape(self.d_model, QD)
        wk = self.key_proj.reshape(self.d_model, KD)
        wv = self.value_proj.reshape(self.d_model, KD)
        self._wt_qkv = torch.cat([wq, wk, wv], dim=1).t().contiguous().to(BF16)
        self._wt_o = self.out.reshape(QD, self.d_model).t().contiguous().to(BF16)

    def initialize_carry(self, batch_size: int, rngs=None) -> KVCache:
        shape = (batch_size, self.max_seq_length, self.num_kv_heads, self.head_dim)
        return KVCache(torch.zeros(shape, dtype=self.dtype, device=self.device),
                       torch.zeros(shape, dtype=self.dtype, device=self.device))

    def update_kv_cache(self, kv_cache: KVCache, seq_pos, key, value) -> KVCache:
        """attention.py:108-114: write (B,1,Nkv,D) key/value at ring slot seq_pos[0,0] % S.  The cache buffers are
        updated in place (the reference donates them) and returned."""
        pos = int(seq_pos.reshape(-1)[0].item()) % self.max_seq_length
        kv_cache.key[:, pos:pos + key.shape[1]] = key.to(kv_cache.key.dtype)
        kv_cache.value[:, pos:pos + value.shape[1]] = value.to(kv_cache.value.dtype)
        return kv_cache

    def __call__(self, inputs: torch.Tensor, seq_pos: torch.Tensor, kv_cache: Optional[KVCache] = None):
        batch, seq, _ = inputs.shape
        Nq, Nkv, D, S = self.num_heads, self.num_kv_heads, self.head_dim, self.max_seq_length
        x = inputs.to(BF16).contiguous().reshape(batch * seq, self.d_model)
        pos = seq_pos.to(torch.int32)
        if kv_cache is not None:
            if seq != 1:
                raise ValueError("the KV-cache branch decodes one token per call (add_seq_dim, util.py:129-130)")
            pos = (pos.expand(batch, 1) if pos.shape[0] == 1 else pos).contiguous().reshape(-1)
            q, _, _ = ops.qkv_rope(x, self._wt_qkv, pos, self._timescale, Nq, Nkv, D, k=kv_cache.key, v=kv_cache.value,
                                   k_row_stride=S * Nkv * D, v_row_stride=S * Nkv * D, cache_S=S)
            o = ops.attn_decode(q, kv_cache.key, kv_cache.value, pos, Nq, Nkv, D)
        else:
            if S < seq:
                raise NotImplementedError("sliding-window training (attention.py:152-161) is unreachable from train_run")
            pos = (pos.expand(batch, seq) if pos.shape[0] == 1 else pos).contiguous().reshape(-1)
            QD, KD = Nq * D, Nkv * D
            qkv = torch.empty(batch * seq, QD + 2 * KD, dtype=BF16, device=x.device)
            ops.qkv_rope(x, self._wt_qkv, pos, self._timescale, Nq, Nkv, D, q=qkv, k=qkv[:, QD:], v=qkv[:, QD + KD:],
                         q_row_stride=QD + 2 * KD, k_row_stride=QD + 2 * KD, v_row_stride=QD + 2 * KD)
            o, _ = ops.attn_train_fwd(qkv, qkv[:, QD:], qkv[:, QD + KD:], batch, seq, Nq, Nkv, D, QD + 2 * KD,
                                      QD + 2 * KD, QD + 2 * KD)
        out = ops.linear(o, self._wt_o)
        return out.reshape(batch, seq, self.d_model), kv_cache
