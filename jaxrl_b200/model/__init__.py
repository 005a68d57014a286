"""Host-side mirror of mapox_trainer/model/* over the B200 kernels (same class / method names and argument meaning)."""

from .attention import AttentionBlock, KVCache  # noqa: F401
from .feed_forward import GLUBlock  # noqa: F401
from .positional_embeddings import apply_rope  # noqa: F401
