"""Synthetic workloads of the BASELINE configs: fixed random Skip-Masks at tile
granularity (C2 / the direct-sparsity variant of C4) and a video-like Q/K generator whose attention maps
are local in (frame, y, x) and temporally coherent across diffusion steps (C3 / C4)."""
import math

import torch

BLOCK_M, BLOCK_N = 128, 176


def tile_counts(seqlen_q, seqlen_k=None):
    seqlen_k = seqlen_q if seqlen_k is None else seqlen_k
    return (seqlen_q + BLOCK_M - 1) // BLOCK_M, (seqlen_k + BLOCK_N - 1) // BLOCK_N


def encode_keep_mask(keep: torch.Tensor) -> torch.Tensor:
    """keep: bool [..., ktiles] (True = compute tile) -> int32 run-length list [..., ktiles+1] in the reference
    format [len, s0, e0, s1, e1, ...] (descending inclusive ranges).  Vectorised (runs on CPU or GPU).
    Rows whose encoding would not fit (more than ktiles entries) raise."""
    kt = keep.shape[-1]
    flat = keep.reshape(-1, kt)
    rows = flat.shape[0]
    rev = flat.flip(-1)                                    # position j <-> tile kt-1-j (descending order)
    prev = torch.cat([torch.zeros_like(rev[:, :1]), rev[:, :-1]], dim=1)
    nxt = torch.cat([rev[:, 1:], torch.zeros_like(rev[:, :1])], dim=1)
    is_start = rev & ~prev
    is_end = rev & ~nxt
    tile = (kt - 1 - torch.arange(kt, device=keep.device)).expand(rows, kt)
    nranges = is_start.sum(dim=1)
    if int(nranges.max()) * 2 > kt:
        raise ValueError("keep mask alternates too finely to fit the ktiles+1 row format")
    out = torch.zeros(rows, kt + 1, dtype=torch.int32, device=keep.device)
    out[:, 0] = (2 * nranges).to(torch.int32)
    rank_s = torch.cumsum(is_start, dim=1) - 1             # index of the range each start belongs to
    rank_e = torch.cumsum(is_end, dim=1) - 1
    r_idx = torch.arange(rows, device=keep.device).unsqueeze(1).expand(rows, kt)
    out[r_idx[is_start], (1 + 2 * rank_s)[is_start]] = tile[is_start].to(torch.int32)
    out[r_idx[is_end], (2 + 2 * rank_e)[is_end]] = tile[is_end].to(torch.int32)
    return out.reshape(*keep.shape[:-1], kt + 1)


def random_skip_list(b, h, qtiles, ktiles, sparsity, seed=1234, device="cpu", run=1):
    """Fixed random Skip-Mask: each (b, h, q-tile) row keeps a Bernoulli(1 - sparsity) subset of K tiles, drawn in
    runs of `run` tiles, last (ragged) tile forced kept.  Returns (list int32 [b,h,qtiles,ktiles+1], keep bool)."""
    g = torch.Generator().manual_seed(seed)
    nrun = (ktiles + run - 1) // run
    keep = torch.rand(b, h, qtiles, nrun, generator=g) >= sparsity
    keep = keep.repeat_interleave(run, dim=-1)[..., :ktiles].contiguous()
    keep[..., ktiles - 1] = True
    keep = keep.to(device)
    return encode_keep_mask(keep), keep


def exact_sparsity_list(b, h, qtiles, ktiles, sparsity, seed=1234, device="cpu", run=4):
    """Like random_skip_list but every row skips exactly round(sparsity * ktiles) tiles (in runs of `run`), so a
    single timed call sits at the target sparsity (the "direct variant" of config C4)."""
    g = torch.Generator().manual_seed(seed)
    nrun = (ktiles - 1 + run - 1) // run                   # the last tile is always kept
    n_skip_runs = min(nrun, int(round(sparsity * ktiles / run)))
    scores = torch.rand(b, h, qtiles, nrun, generator=g)
    kth = scores.sort(dim=-1).values[..., n_skip_runs - 1:n_skip_runs] if n_skip_runs > 0 else None
    skip = (scores <= kth) if kth is not None else torch.zeros_like(scores, dtype=torch.bool)
    keep = ~skip.repeat_interleave(run, dim=-1)[..., :ktiles - 1]
    keep = torch.cat([keep, torch.ones_like(keep[..., :1])], dim=-1).contiguous().to(device)
    return encode_keep_mask(keep), keep


class VideoLikeQKV:
    """Video-like Q/K/V: token t <-> (f, y, x) on a frames x height x width grid; q and k share a fixed random-Fourier
    positional embedding E (inner product decays with grid distance) plus AR(1) noise across diffusion steps.
        q = a_h * E + sigma * eps_q(step),   k = a_h * E + sigma * eps_k(step),   v = randn
    a_h varies per head in [0.5, 2] * amp so different heads reach different sparsities."""

    def __init__(self, batch, heads, grid=(21, 45, 80), head_dim=128, amp=14.0, sigma=1.0, rho=0.95, seed=0,
                 bandwidth=(0.9, 0.35, 0.35), device="cuda", seq_len=None):
        f, y, x = grid
        self.S = f * y * x if seq_len is None else seq_len
        self.B, self.H, self.D = batch, heads, head_dim
        self.sigma, self.rho, self.device = sigma, rho, device
        g = torch.Generator(device="cpu").manual_seed(7)
        t = torch.arange(self.S)
        pos = torch.stack([t // (y * x), (t // x) % y, t % x], dim=1).float()
        W = torch.randn(3, head_dim, generator=g) * torch.tensor(bandwidth).unsqueeze(1)
        phase = torch.rand(head_dim, generator=g) * 2 * math.pi
        E = math.sqrt(2.0 / head_dim) * torch.cos(pos @ W + phase)            # (S, D), |E_t| ~ 1
        gh = torch.Generator(device="cpu").manual_seed(seed + 100)
        a_h = amp * (0.5 + 1.5 * torch.rand(heads, generator=gh))             # per-head amplitude
        self.base = (E.unsqueeze(1) * a_h.view(1, heads, 1)).to(device)       # (S, H, D)
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.eps_q = self._noise()
        self.eps_k = self._noise()
        self.step = 0

    def _noise(self):
        return torch.randn(self.B, self.S, self.H, self.D, generator=self.gen, device=self.device)

    def next(self):
        """Q, K, V (B, S, H, D) bf16 for the next diffusion step."""
        if self.step > 0:
            c = math.sqrt(1 - self.rho ** 2)
            self.eps_q.mul_(self.rho).add_(self._noise(), alpha=c)
            self.eps_k.mul_(self.rho).add_(self._noise(), alpha=c)
        self.step += 1
        q = (self.base.unsqueeze(0) + self.sigma * self.eps_q).to(torch.bfloat16)
        k = (self.base.unsqueeze(0) + self.sigma * self.eps_k).to(torch.bfloat16)
        v = torch.randn(self.B, self.S, self.H, self.D, generator=self.gen, device=self.device).to(torch.bfloat16)
        return q, k, v


def flops_dense(b, h, sq, sk, d):
    """4*B*H*Sq*Sk*D (hopper/_internal/benchmarks/benchmark_attn.py:62-73 of the reference with d == dv)."""
    return 4.0 * b * h * sq * sk * d


def flops_executed(read_list, seqlen_q, seqlen_k, d, batch=None):
    """FLOPs of the tiles actually listed: 4*D * sum_rows rows(m) * sum_{n in list} cols(n)."""
    rl = read_list if batch is None else read_list[:batch]
    rl = rl.to(torch.int64)
    b, h, qt, kp1 = rl.shape
    kt = kp1 - 1
    ln = rl[..., 0].clamp(0, kt)
    ent = rl[..., 1:]
    pos = torch.arange(kt, device=rl.device)
    valid = pos < ln.unsqueeze(-1)
    is_start = (pos % 2 == 0) & valid
    is_end = (pos % 2 == 1) & valid
    # cols covered by range [e, s] = min((s+1)*BN, Sk) - e*BN
    hi = torch.clamp((ent + 1) * BLOCK_N, max=seqlen_k)
    lo = ent * BLOCK_N
    cols = (hi * is_start).sum(-1) - (lo * is_end).sum(-1)                    # [b,h,qt]
    rows = torch.clamp(seqlen_q - torch.arange(qt, device=rl.device) * BLOCK_M, max=BLOCK_M)
    return 4.0 * d * float((cols * rows.view(1, 1, qt)).sum().item())


def update_bytes(read_list, write_list=None, batch=None):
    """Algorithmic HBM bytes of one la_skip_update call: every row reads its list (len + 1 words) and
    the statistic of each tile it visited, and writes the next list (len' + 1 words)."""
    rl = (read_list if batch is None else read_list[:batch]).to(torch.int64)
    wl = rl if write_list is None else (write_list if batch is None else write_list[:batch]).to(torch.int64)
    kt = rl.shape[-1] - 1
    ln = rl[..., 0].clamp(0, kt)
    ent = rl[..., 1:]
    pos = torch.arange(kt, device=rl.device)
    valid = pos < ln.unsqueeze(-1)
    visited = ((ent + 1) * ((pos % 2 == 0) & valid)).sum(-1) - (ent * ((pos % 2 == 1) & valid)).sum(-1)
    words = (ln + 1).sum() + visited.sum() + (wl[..., 0].clamp(0, kt) + 1).sum()
    return 4.0 * float(words.item())
