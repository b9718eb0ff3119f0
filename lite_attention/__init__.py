"""Drop-in alias: `from lite_attention import LiteAttention, SeqParallelLiteAttention` resolves to the B200
implementation in liteattention_b200 (same names as hopper/__init__.py:1-6 of the reference)."""
from liteattention_b200 import __version__  # noqa: F401
from liteattention_b200.lite_attention import LiteAttention, SeqParallelLiteAttention  # noqa: F401

__all__ = ["LiteAttention", "SeqParallelLiteAttention"]
