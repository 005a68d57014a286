"""jaxrl_b200 - B200 (sm_100a) hot path of mapox_trainer's transformer-over-time policy.

Host-side mirror of the reference's module API over a C-ABI CUDA library (libmapox_b200.so).  The library is loaded
by `jaxrl_b200._lib` the first time a compute module (`ops`, `engine`, `trainer`, `model.*`, `values`, `rollout`) is
imported and that import fails loudly when the .so is missing - there is no CPU fallback.  `config`, `params` and
`dist` are plain host logic and import without it.
"""

from .config import ModelConfig, PPOConfig, OptimizerConfig, RunConfig, NAMED, from_reference_json  # noqa: F401

__version__ = "0.2.0"
