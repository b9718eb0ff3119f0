"""Fused 3-D RoPE + bf16 cast for Q / K: the step the reference's Wan integration
runs in front of every LiteAttention call (reference README.md:301-315),

    q_rope = rope_apply(q, grid_sizes, freqs).bfloat16()

as ONE HBM pass (`la_rope_cast_sm100`, liteattention_b200/csrc/la_rope_cast.cu) instead of the float64 complex
round trip of Wan2.1's `rope_apply` (wan/modules/model.py).  Same arguments as that function."""
import weakref
from typing import Dict, Tuple

import torch

from . import _native

# (id of the freqs tensor, device) -> (weakref to that tensor, its _version when the table was built, table).  Keyed on
# the tensor OBJECT, not its address: a freed-and-reallocated or in-place modified `freqs` never meets a stale table,
# and entries die with their tensors.
_TABLES: Dict[Tuple[int, torch.device], tuple] = {}
# id of a grid_sizes tensor -> (weakref, _version, max grid extent): avoids a device sync per call when grid_sizes lives on the GPU
_GRID_MAX: Dict[int, tuple] = {}


def _cos_sin_table(freqs: torch.Tensor, device: torch.device) -> torch.Tensor:
    """freqs: complex [max_pos, d/2] as built by the Wan model (frames | height | width tables concatenated along
    dim 1) -> fp32 [max_pos, d/2, 2] (cos, sin) on `device`, cached per (tensor object, version, device)."""
    key = (id(freqs), device)
    ent = _TABLES.get(key)
    if ent is not None and ent[0]() is freqs and ent[1] == freqs._version:
        return ent[2]
    f = freqs.to(torch.complex128)
    t = torch.stack([f.real, f.imag], dim=-1).to(torch.float32).to(device).contiguous()
    _TABLES[key] = (weakref.ref(freqs, lambda _r, k=key: _TABLES.pop(k, None)), freqs._version, t)
    return t


def _grid_max(grid_sizes: torch.Tensor) -> int:
    key = id(grid_sizes)
    ent = _GRID_MAX.get(key)
    if ent is not None and ent[0]() is grid_sizes and ent[1] == grid_sizes._version:
        return ent[2]
    m = int(grid_sizes.max().item()) if grid_sizes.numel() else 0
    _GRID_MAX[key] = (weakref.ref(grid_sizes, lambda _r, k=key: _GRID_MAX.pop(k, None)), grid_sizes._version, m)
    return m


def rope_apply_bf16(x: torch.Tensor, grid_sizes: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """x: (batch, seq_len, heads, head_dim) fp32 or bf16 on a CUDA device; grid_sizes: int [batch, 3] (frames, height,
    width); freqs: complex [max_pos, head_dim/2].  Returns bf16 (batch, seq_len, heads, head_dim) =
    rope_apply(x, grid_sizes, freqs).bfloat16() of the Wan model."""
    if not x.is_cuda:
        raise RuntimeError("rope_apply_bf16: CUDA tensors only (liteattention_b200 has no CPU path)")
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise NotImplementedError(f"rope_apply_bf16: dtype {x.dtype} is not supported (fp32 or bf16)")
    b, s, h, d = x.shape
    if d % 8 != 0 or freqs.shape[1] != d // 2:
        raise ValueError("rope_apply_bf16: head_dim must be a multiple of 8 and freqs must be [max_pos, head_dim/2]")
    elt = x.element_size()
    if x.stride(-1) != 1 or x.data_ptr() % 16 or any((st * elt) % 16 for st in x.stride()[:3]):
        x = x.contiguous()
    grid = grid_sizes.to(device=x.device, dtype=torch.int32).contiguous()
    if _grid_max(grid_sizes) > freqs.shape[0]:
        raise ValueError("rope_apply_bf16: a grid dimension exceeds the position table")
    out = torch.empty((b, s, h, d), dtype=torch.bfloat16, device=x.device)
    _native.rope_cast(x, out, _cos_sin_table(freqs, x.device), grid)
    return out
