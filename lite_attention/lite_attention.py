from liteattention_b200.lite_attention import *  # noqa: F401,F403
from liteattention_b200.lite_attention import LiteAttention, SeqParallelLiteAttention  # noqa: F401
