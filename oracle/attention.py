"""Tile-exact CPU restatement (torch, fp32) of the reference's skippable attention forward.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, per (batch, head, 128-row Q tile):
  first visited tile  : mainloop_fwd_sm90_tma_gmma_ws.hpp:1612-1663  (seqlen mask mask.h:66-76, never skip-tested)
  every further tile  : mainloop :1667-1755; softmax.h:139-222 (max, rescale, skip predicate :194),
                        softmax.h:81-121 (P = exp2(S*c - m*c)), :263-273 (row sum in fp32 before bf16 rounding)
  P -> bf16 (RN) before PV: utils.h:211-225;  fp32 accumulate
  finalize            : softmax.h:275-296 (O / l, LSE = m*scale + ln l);  epilogue_fwd.hpp:241-242, 298-330
  list codec          : oracle/skiplist.py
TMA zero-fill: Q rows >= S and K/V rows >= S are zeros (only the FIRST visited tile gets -inf on cols >= S).
"""
import math

import torch

from . import skiplist as sl

LOG2E = 1.4426950408889634


class OnlineSoftmaxState:
    """The reference's per-Q-tile softmax state machine, one call per visited K tile (all 128 rows at once).

    step()      = flash::Softmax::max_get_scale_detect_qk_skip (softmax.h:139-222: local max, running max, alpha,
                  row_sum *= alpha, skip predicate :194 with m_prev = the running max BEFORE this tile, vote :207-218)
                  followed by online_softmax (:263-273: P = exp2(S*c - m*c), scale_apply_exp2 :81-121, row sum of the
                  fp32 P BEFORE bf16 rounding) and the bf16 conversion of P (utils.h:211-225, RN).
    finalize()  = flash::Softmax::finalize (:275-296).
    Pinned on the GPU box against the reference's own code: oracle/ref_softmax_harness.cu compiles softmax.h / mask.h
    from /root/reference and tests/test_ref_softmax_gpu.py compares every quantity returned here with it."""

    def __init__(self, c, scale, thr, rows=128):
        self.c = torch.as_tensor(c, dtype=torch.float32)
        self.scale = float(scale)
        self.thr = torch.tensor(thr, dtype=torch.float32)
        self.m_run = torch.full((rows,), float("-inf"))
        self.l = torch.zeros(rows)
        self.first = True

    def step(self, S_):
        """S_: (rows, cols) fp32 raw scores, already masked if this is the first tile.
        Returns (alpha (rows,), P fp32 (rows, cols), P_bf16, vote (bool or None for the first tile), stat (float))."""
        c = self.c
        m_loc = S_.max(dim=1).values
        m_new = torch.maximum(self.m_run, m_loc)
        if not self.first:
            d = (m_loc - self.m_run) * c
            do_rows = d > self.thr                       # NaN compares false (softmax.h:194)
            vote = not bool(do_rows.any())               # skip = !any(do_qk) (:207), AND over the 8 warps (:1721-1725)
            dd = torch.where(torch.isnan(d), torch.full_like(d, float("-inf")), d)
            stat = float(dd.max())
            m_safe = m_new                               # Check_inf = false on further tiles (mainloop :1810)
        else:
            vote, stat = None, float("inf")
            # Check_inf = true on the first tile (:1632, :1641): a fully masked row uses 0 instead of -inf
            m_safe = torch.where(torch.isinf(m_new) & (m_new < 0), torch.zeros_like(m_new), m_new)
        alpha = torch.ones_like(m_new) if self.first else torch.exp2((self.m_run - m_safe) * c)
        P = torch.exp2(S_ * c - (m_safe * c)[:, None])
        self.l = (self.l * alpha + P.sum(dim=1)) if not self.first else P.sum(dim=1)
        self.m_run = m_new
        self.first = False
        return alpha, P, P.to(torch.bfloat16), vote, stat

    def finalize(self):
        """(inv (rows,), lse (rows,)): softmax.h:283-293."""
        l = self.l
        bad = (l == 0) | torch.isnan(l)
        inv = torch.where(bad, torch.zeros_like(l), 1.0 / l)
        # row_max * (softmax_scale_log2 * ln 2) + log(sum)  (:293)
        lse = torch.where(bad, torch.full_like(l, float("-inf")),
                          self.m_run * (self.c * torch.tensor(math.log(2.0), dtype=torch.float32)) + torch.log(l))
        return inv, lse


def softmax_tile_sequence(S_tiles, c, scale, thr, first_tile_valid_cols=None):
    """Run OnlineSoftmaxState over raw-score tiles in visit order.  first_tile_valid_cols: columns >= this of the
    FIRST tile are out of range and get -inf (mask.h:66-76, applied to that tile only, mainloop :1626).
    Returns dict(tiles=[dict(m_run, alpha, p_bf16, vote, stat)], inv, lse)."""
    st = OnlineSoftmaxState(c, scale, thr, rows=S_tiles[0].shape[0])
    res = []
    for i, S_ in enumerate(S_tiles):
        S_ = S_.clone().float()
        if i == 0 and first_tile_valid_cols is not None:
            S_[:, first_tile_valid_cols:] = float("-inf")
        alpha, P, Pb, vote, stat = st.step(S_)
        res.append(dict(m_run=st.m_run.clone(), alpha=alpha, p_bf16=Pb, vote=vote, stat=stat))
    inv, lse = st.finalize()
    return dict(tiles=res, inv=inv, lse=lse)


def lite_attention_oracle(q, k, v, softmax_scale=None, read_list=None, must_do_list=None, thr=-3.0,
                          on_overflow="copy", q_tiles=None):
    """q,k,v: (B,S,H,D) bf16 (CPU).  read_list / must_do_list: int32 [>=B,H,qtiles,ktiles+1] or None (dense).
    Returns dict(out bf16 (B,Sq,H,D), out_f32 (the same before the final bf16 rounding), lse fp32 (B,H,Sq), write_list int32 like read_list (or None),
                 stat fp32 (B,H,qtiles,ktiles) with NaN at unvisited tiles and +inf at each row's first tile,
                 visited = number of (q-tile,k-tile) pairs computed).
    q_tiles: optional iterable of Q-tile indices; only those rows are computed (spot checks at sizes the full walk
    would not finish in seconds), everything else keeps its initial value."""
    B, Sq, H, D = q.shape
    Sk, Hk = k.shape[1], k.shape[2]
    assert D == 128 and q.dtype == torch.bfloat16
    bm, bn = sl.get_MN(D, 2)
    qtiles, ktiles = sl.ceil_div(Sq, bm), sl.ceil_div(Sk, bn)
    scale = D ** -0.5 if softmax_scale is None else float(softmax_scale)
    c = torch.tensor(scale, dtype=torch.float32) * torch.tensor(LOG2E, dtype=torch.float32)  # fp32 product, mainloop :760
    thr32 = torch.tensor(thr, dtype=torch.float32)

    qf = torch.zeros(B, qtiles * bm, H, D)
    qf[:, :Sq] = q.float()
    kf = torch.zeros(B, ktiles * bn, Hk, D)
    kf[:, :Sk] = k.float()
    vf = torch.zeros(B, ktiles * bn, Hk, D)
    vf[:, :Sk] = v.float()

    out = torch.zeros(B, Sq, H, D, dtype=torch.bfloat16)
    out_f32 = torch.zeros(B, Sq, H, D)
    lse = torch.full((B, H, Sq), float("-inf"))
    stat = torch.full((B, H, qtiles, ktiles), float("nan"))
    write_list = None if read_list is None else read_list.clone()
    n_visited = 0
    rep = H // Hk
    for b in range(B):
        for h in range(H):
            hk = h // rep
            for m in (range(qtiles) if q_tiles is None else q_tiles):
                rd = sl.init_row(ktiles) if read_list is None else read_list[b, h, m].tolist()
                Qt = qf[b, m * bm:(m + 1) * bm, h]
                tiles = sl.visited_tiles(rd, ktiles)
                votes = {}
                st = OnlineSoftmaxState(c, scale, thr, rows=bm)
                O = torch.zeros(bm, D)
                for idx, n in enumerate(tiles):
                    Kt = kf[b, n * bn:(n + 1) * bn, hk]
                    Vt = vf[b, n * bn:(n + 1) * bn, hk]
                    S_ = Qt @ Kt.T
                    if idx == 0:
                        col = torch.arange(n * bn, (n + 1) * bn)
                        S_[:, col >= Sk] = float("-inf")
                    alpha, P, Pb, vote, st_n = st.step(S_)
                    if idx > 0:
                        votes[n] = vote
                    stat[b, h, m, n] = st_n
                    O = O * alpha[:, None] + Pb.float() @ Vt
                    n_visited += 1
                if tiles:
                    inv, ls = st.finalize()
                    rows = min(bm, Sq - m * bm)
                    out_f32[b, m * bm:m * bm + rows, h] = (O * inv[:, None])[:rows]
                    out[b, m * bm:m * bm + rows, h] = (O * inv[:, None])[:rows].to(torch.bfloat16)
                    lse[b, h, m * bm:m * bm + rows] = ls[:rows]
                if write_list is not None:
                    md = None if must_do_list is None else must_do_list[b, h, m].tolist()
                    row, _ = sl.skip_list_step(rd, lambda n: votes[n], md, ktiles, on_overflow=on_overflow)
                    write_list[b, h, m, :len(row)] = torch.tensor(row, dtype=torch.int32)
    return dict(out=out, out_f32=out_f32, lse=lse, write_list=write_list, stat=stat, visited=n_visited)


def dense_attention_ref(q, k, v, softmax_scale=None):
    """fp32 reference of plain softmax attention (same math as hopper/tests/test_util.py:226-348
    attention_ref without masks): q,k,v (B,S,H,D) any float dtype -> (out fp32 (B,S,H,D), lse fp32 (B,H,S))."""
    D = q.shape[-1]
    scale = D ** -0.5 if softmax_scale is None else float(softmax_scale)
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    if kf.shape[1] != qf.shape[1]:
        rep = qf.shape[1] // kf.shape[1]
        kf, vf = kf.repeat_interleave(rep, 1), vf.repeat_interleave(rep, 1)
    s = (qf @ kf.transpose(-1, -2)) * scale
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ vf
    return o.permute(0, 2, 1, 3).contiguous(), lse
