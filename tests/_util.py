"""Shared helpers for the GPU parity tests."""

import os

import torch

# Flash-style (online-softmax) attention rounds the UN-normalised probabilities to bf16 - as cuDNN's fused SDPA, the
# reference's GPU implementation, does - while the XLA restatement rounds after normalising: outputs may differ by up
# to 2 bf16 ulps (2 * 2^-7 relative at the bottom of a binade).  Every other bf16 result is held to 1e-2.
ULP2 = 2.0 ** -6


def log(msg):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.log", "a") as f:
        f.write(msg + "\n")


def assert_close_bf16(got, ref, name, rtol=1e-2, abs_floor_frac=4e-3, allow_frac=0.0, max_err_frac=None):
    """|got-ref| <= rtol*|ref| + abs_floor_frac*rms(ref) elementwise (bf16 results).

    allow_frac: fraction of elements allowed outside that band (bf16 rounding flips upstream of a deep stack);
    max_err_frac: hard cap on the largest error as a fraction of rms(ref) when allow_frac > 0."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite output"
    rms = ref.pow(2).mean().sqrt().item() + 1e-12
    err = (got - ref).abs()
    tol = rtol * ref.abs() + abs_floor_frac * rms
    bad = (err > tol).float().mean().item()
    log(f"{name}: max_abs_err={err.max().item():.4e} rms_ref={rms:.4e} frac_bad={bad:.3e}")
    assert bad <= allow_frac, f"{name}: {bad * 100:.4f}% elements out of tolerance, max err {err.max().item():.4e} (rms {rms:.3e})"
    if allow_frac > 0 and max_err_frac is not None:
        assert err.max().item() <= max_err_frac * rms, f"{name}: max err {err.max().item():.4e} > {max_err_frac} * rms {rms:.3e}"


def rope_timescale(D, base=10_000.0):
    return torch.tensor(base, dtype=torch.float32) ** (2.0 * torch.arange(0, D // 2, dtype=torch.float32) / D)
