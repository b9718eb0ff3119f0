"""Threshold calibration (the reference describes "error-calibrated thresholds",
README.md:14, and publishes no code for it).

Because this implementation keeps the per-tile QK-skip statistic as a tensor (the reference folds it into a vote
inside its fused kernel), a threshold can be chosen AFTER a forward pass without recomputing attention: one forward
writes the statistic of every visited tile, then the run-length update kernel -- 60 microseconds at the Wan2.1-14B
shape -- is replayed under bisection until the written list reaches a target sparsity."""
from typing import Optional, Tuple

import torch

from . import _native
from .lite_attention import LiteAttention


def sparsity_at_threshold(read_list: torch.Tensor, tile_stat: torch.Tensor, thr: float,
                          must_do_list: Optional[torch.Tensor] = None,
                          scratch: Optional[torch.Tensor] = None) -> Tuple[float, torch.Tensor]:
    """Sparsity of the list la_skip_update would write from (read_list, tile_stat) at threshold thr."""
    b, h, qtiles, ktiles = tile_stat.shape
    wl = torch.zeros_like(read_list[:b]) if scratch is None else scratch
    _native.skip_update(read_list, must_do_list, wl, tile_stat, b, h, qtiles, ktiles, float(thr))
    return LiteAttention.sparsity(wl[:b]), wl


def calibrate_threshold(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, target_sparsity: float,
                        read_list: Optional[torch.Tensor] = None, scale: Optional[float] = None,
                        must_do_list: Optional[torch.Tensor] = None, lo: float = -64.0, hi: float = -1e-3,
                        iters: int = 24) -> Tuple[float, float, torch.Tensor]:
    """Find thr in [lo, hi] (exp2 domain, negative) such that ONE QK-skip step from `read_list` (default: the dense
    initial list) on these q, k, v leaves a list of `target_sparsity`.  Returns (thr, achieved sparsity, list).
    Sparsity is non-decreasing in thr (a larger threshold can only turn "do" votes into "skip"), so plain bisection
    works; if even `hi` does not reach the target the result at `hi` is returned."""
    b, s, h, d = q.shape
    if read_list is None:
        read_list = LiteAttention.init_skip_list(b, s, h, d, False, q.dtype, q.device)[0]
    qtiles, ktiles = read_list.shape[-2], read_list.shape[-1] - 1
    out = torch.empty_like(q)
    lse = torch.empty(b, h, s, device=q.device, dtype=torch.float32)
    stat = torch.empty(b, h, qtiles, ktiles, device=q.device, dtype=torch.float32)
    _native.fwd(q, k, v, out, lse, d ** -0.5 if scale is None else scale, read_list, stat)
    scratch = torch.zeros_like(read_list[:b])
    sp_hi, wl = sparsity_at_threshold(read_list, stat, hi, must_do_list, scratch)
    if sp_hi <= target_sparsity:
        return hi, sp_hi, wl.clone()
    best = (hi, sp_hi)
    a, c = lo, hi
    for _ in range(iters):
        m = 0.5 * (a + c)
        sp, _ = sparsity_at_threshold(read_list, stat, m, must_do_list, scratch)
        if sp < target_sparsity:
            a = m
        else:
            c = m
            best = (m, sp)
    sp, wl = sparsity_at_threshold(read_list, stat, best[0], must_do_list, scratch)
    return best[0], sp, wl.clone()
