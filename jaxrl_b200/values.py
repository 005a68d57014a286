"""Mirror of mapox_trainer/values.py (HL-Gauss value head)."""

from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


def calculate_supports(config):
    """values.py:12-19 -> (support (1, n+1), centers (n,)) fp32 (host tensors)."""
    support = torch.linspace(config.min, config.max, config.n_logits + 1, dtype=F32)
    centers = (support[:-1] + support[1:]) / 2
    return support[None, :], centers


class HlGaussValue:
    def __init__(self, in_features: int, hl_gauss_config, *, rngs=0, device="cuda"):
        self._min, self._max, self._sigma = hl_gauss_config.min, hl_gauss_config.max, hl_gauss_config.sigma
        self.n_logits = hl_gauss_config.n_logits
        gen = torch.Generator().manual_seed(int(rngs))
        dev = torch.device(device)
        self.kernel = (torch.randn(in_features, self.n_logits, generator=gen) * (1.0 / in_features) ** 0.5).to(dev)
        self.bias = torch.zeros(self.n_logits, device=dev)
        supports, centers = calculate_supports(hl_gauss_config)
        self._supports = supports.to(dev)
        self._centers = centers.to(dev)
        self.refresh()

    def refresh(self):
        K, N = self.kernel.shape
        if N > 64:
            raise ValueError("n_logits > 64 is not supported")
        wt = torch.zeros(128, K, dtype=BF16, device=self.kernel.device)
        hi = self.kernel.t().to(BF16)
        wt[:N] = hi
        wt[64:64 + N] = (self.kernel.t() - hi.float()).to(BF16)
        self._wt_hilo = wt

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 Linear on f32(bf16 x) (values.py:34,40-41)."""
        lead = x.shape[:-1]
        xb = x.to(BF16).contiguous().reshape(-1, x.shape[-1])
        out = torch.empty(xb.shape[0], self.n_logits, dtype=F32, device=x.device)
        ops.linear_f32w(xb, self._wt_hilo, self.bias, out, self.n_logits)
        return out.reshape(*lead, self.n_logits)

    def get_value(self, logits: torch.Tensor) -> torch.Tensor:
        return ops.hlgauss_value(logits.to(F32).contiguous(), self._centers)

    def get_loss(self, logits: torch.Tensor, target_values: torch.Tensor) -> torch.Tensor:
        """Soft-label cross entropy, mean over (b t) (values.py:47-63) -> 0-d tensor."""
        loss, _ = ops.hlgauss_loss(logits.to(F32).contiguous(), target_values.to(F32).contiguous().reshape(-1),
                                   self._supports.reshape(-1), self._min, self._max, self._sigma, want_grad=False)
        return loss.reshape(())

    def get_loss_and_grad(self, logits, target_values, grad_scale: float = 1.0):
        loss, dl = ops.hlgauss_loss(logits.to(F32).contiguous(), target_values.to(F32).contiguous().reshape(-1),
                                    self._supports.reshape(-1), self._min, self._max, self._sigma, grad_scale)
        return loss.reshape(()), dl


def HlGaussConfig(min: float, max: float, n_logits: int, sigma: float, type: str = "hl_gauss"):
    return SimpleNamespace(type=type, min=min, max=max, n_logits=n_logits, sigma=sigma)
