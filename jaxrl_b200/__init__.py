"""jaxrl_b200 - B200 (sm_100a) hot path of mapox_trainer's transformer-over-time policy.

Host-side mirror of the reference's module API over a C-ABI CUDA library (libmapox_b200.so).
Importing this package loads the shared library; there is no CPU fallback.
"""

from . import _lib  # noqa: F401  (fails loudly if the CUDA extension is missing)
from .config import ModelConfig, PPOConfig, OptimizerConfig, RunConfig, NAMED, from_reference_json  # noqa: F401

__version__ = "0.1.0"
