"""CPU oracle for the mapox_trainer transformer-over-time policy hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``jaxrl_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline - never as the product path.

Parity status: the reference (gabe00122/jaxrl, ``mapox_trainer``) is pure
JAX/flax and cannot be imported in this image (no jax/flax/mapox wheels), so
this is a *restatement* in torch-CPU fp32 with explicit bf16 rounding points.
It is pinned only by the known-answer values and properties held in the
reference's own tests (tests/test_attention.py, test_rope.py,
test_advantage.py, test_values.py - re-expressed in tests/test_oracle_*.py).
Everything those tests do not cover (GLU, RMSNorm, trunk, tied logits,
ppo_loss, every bf16 result) is **parity unpinned**.
"""
