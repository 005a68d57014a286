"""Shared helpers for the GPU parity tests."""

import os

import torch


def log(msg):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.log", "a") as f:
        f.write(msg + "\n")


def assert_close_bf16(got, ref, name, rtol=1e-2, abs_floor_frac=4e-3):
    """|got-ref| <= rtol*|ref| + abs_floor_frac*rms(ref) elementwise (bf16 results)."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite output"
    rms = ref.pow(2).mean().sqrt().item() + 1e-12
    err = (got - ref).abs()
    tol = rtol * ref.abs() + abs_floor_frac * rms
    bad = (err > tol).float().mean().item()
    log(f"{name}: max_abs_err={err.max().item():.4e} rms_ref={rms:.4e} frac_bad={bad:.3e}")
    assert bad == 0.0, f"{name}: {bad * 100:.4f}% elements out of tolerance, max err {err.max().item():.4e} (rms {rms:.3e})"


def rope_timescale(D, base=10_000.0):
    return torch.tensor(base, dtype=torch.float32) ** (2.0 * torch.arange(0, D // 2, dtype=torch.float32) / D)
