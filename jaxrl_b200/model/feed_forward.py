"""Mirror of mapox_trainer/model/feed_forward.py (GLUBlock)."""

from __future__ import annotations

import torch

from .. import ops

BF16 = torch.bfloat16


class GLUBlock:
    def __init__(self, d_model: int, hidden_features: int, activation="gelu", *, kernel_init=None, dtype=BF16,
                 param_dtype=torch.float32, rngs=0, device="cuda"):
        if activation not in ("gelu", torch.nn.functional.gelu):
            raise NotImplementedError("only gelu (tanh approximation, jax.nn.gelu default) is on the B200 path")
        if hidden_features % 128 != 0:
            raise ValueError("hidden_features must be a multiple of 128")
        self.d_model, self.hidden_features = d_model, hidden_features
        gen = torch.Generator().manual_seed(int(rngs))
        init = kernel_init if kernel_init is not None else (
            lambda shape: (torch.rand(shape, generator=gen) * 2 - 1) * (6.0 / (shape[0] + shape[1])) ** 0.5)
        dev = torch.device(device)
        self.up_proj = init((d_model, hidden_features)).to(dev)
        self.up_gate = init((d_model, hidden_features)).to(dev)
        self.down_proj = init((hidden_features, d_model)).to(dev)
        self.refresh()

    def refresh(self):
        self._wt_ug = ops.pack_glu_weight(self.up_proj, self.up_gate)
        self._wt_d = self.down_proj.t().contiguous().to(BF16)

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        shape = inputs.shape
        x = inputs.to(BF16).contiguous().reshape(-1, self.d_model)
        h = ops.glu_up(x, self._wt_ug, self.hidden_features)
        return ops.linear(h, self._wt_d).reshape(shape)
