"""Mirror of mapox_trainer/model/positional_embeddings.py."""

from __future__ import annotations

import torch

from .. import ops

_MAX_WAVELENGTH = 10_000


def rope_timescale(head_dim: int, max_wavelength: float, device) -> torch.Tensor:
    frac = 2.0 * torch.arange(0, head_dim // 2, dtype=torch.float32) / head_dim
    return (torch.tensor(float(max_wavelength), dtype=torch.float32) ** frac).to(device)


def apply_rope(inputs: torch.Tensor, positions: torch.Tensor, head_dim: int,
               max_wavelength: float = _MAX_WAVELENGTH) -> torch.Tensor:
    """positional_embeddings.py:7-27.  inputs (B,T,N,D) bf16 CUDA, positions (B|1,T) int -> same shape/dtype."""
    if inputs.dtype != torch.bfloat16:
        raise TypeError("the B200 path computes RoPE on bfloat16 activations (fp32 math, one rounding)")
    B, T, N, D = inputs.shape
    if D != head_dim:
        raise ValueError(f"last dimension {D} != head_dim {head_dim}")
    pos = positions.to(torch.int32)
    if pos.shape[0] == 1 and B > 1:
        pos = pos.expand(B, T)
    pos = pos.contiguous().reshape(-1)
    out = inputs.contiguous().clone()
    ops.rope_(out, N * D, pos, rope_timescale(D, max_wavelength, inputs.device), B * T, N, D, inverse=False)
    return out
