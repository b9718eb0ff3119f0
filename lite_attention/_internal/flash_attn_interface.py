"""Alias of liteattention_b200.flash_attn_interface under the reference's module path
(hopper/_internal/flash_attn_interface.py)."""
import lite_attention._C  # noqa: F401  (registers operators with PyTorch)
from liteattention_b200.flash_attn_interface import (  # noqa: F401
    flash_attn_func, flash_attn_combine, FlashAttnFunc, _flash_attn_forward, maybe_contiguous, flash_attn_3_cuda)
