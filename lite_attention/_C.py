"""`import lite_attention._C` registers torch.ops.lite_attention.* exactly like importing the reference's
extension module does (hopper/_internal/flash_attn_interface.py:10, flash_api.cpp:1722-1824)."""
import liteattention_b200.flash_attn_interface  # noqa: F401
