"""liteattention_b200 -- Blackwell-native (B200, sm_100a) QK-Skip attention forward behind the
moonmath-ai/LiteAttention Python surface (hopper/__init__.py:1-6 of the reference)."""
__version__ = "0.2.0"

from .lite_attention import LiteAttention, SeqParallelLiteAttention  # noqa: E402,F401
from .flash_attn_interface import flash_attn_func, flash_attn_combine  # noqa: E402,F401

# Extensions beyond the reference's surface live in their own modules and are imported on demand:
#   liteattention_b200.rope.rope_apply_bf16            fused 3-D RoPE + bf16 cast (the caller-side step)
#   liteattention_b200.calibrate.calibrate_threshold   threshold for a target sparsity
#   liteattention_b200.dist.BatchParallelLiteAttention / UlyssesLiteAttention   multi-GPU drivers
__all__ = ["LiteAttention", "SeqParallelLiteAttention", "flash_attn_func", "flash_attn_combine"]
