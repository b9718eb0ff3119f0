"""GPU: pins the oracle's floating-point half to the REFERENCE's OWN code.

oracle/_ref/softmax_ref is flash::Softmax / flash::Mask / convert_type_out compiled from
/root/reference/hopper/_internal/cpp/{softmax.h,mask.h,utils.h} (oracle/ref_softmax_harness.cu, `make -C oracle ref`)
and run on the B200 with the FA3 accumulator layout that CuTe itself produces.  For sequences of S tiles it returns
the running row max, the rescale factor alpha, the tile's skip vote, the bf16 P fed to the PV GEMM, and the final
LSE and 1/l.  They are compared with oracle/attention.py (OnlineSoftmaxState) and -- the vote -- with the statistic
la_fwd_kernel emits through the C ABI.  Skipped only when the binary is absent (it is built in the container, where
/root/reference exists, by __graft_entry__.build(), and travels to the GPU box)."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from oracle import attention as oa
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "softmax_ref")
BM, BN = 128, 176

needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/softmax_ref not built (needs /root/reference)")


def run_reference_softmax(S_tiles, n_block_first, seqlen_q, seqlen_k, m_block, scale_log2, thr):
    """S_tiles: float32 array [T, 128, 176] of raw scores in visit order (first tile unmasked)."""
    S_tiles = np.ascontiguousarray(S_tiles, dtype=np.float32)
    T = S_tiles.shape[0]
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<5i2f", T, n_block_first, seqlen_q, seqlen_k, m_block, scale_log2, thr))
            f.write(S_tiles.tobytes())
        subprocess.check_call([REF_BIN, fin, fout])
        raw = np.fromfile(fout, dtype=np.float32)
    per = 128 + 128 + 1 + BM * BN
    assert raw.size == T * per + 256
    tiles = []
    for t in range(T):
        o = raw[t * per:(t + 1) * per]
        tiles.append(dict(m_run=o[:128], alpha=o[128:256], vote=bool(o[256]), p=o[257:].reshape(BM, BN)))
    return dict(tiles=tiles, lse=raw[T * per:T * per + 128], inv=raw[T * per + 128:])


def bf16_ulp_steps(a, b):
    """|a - b| in units of bf16 spacing at max(|a|,|b|) (a, b hold bf16-representable fp32 values)."""
    a = torch.as_tensor(a).float()
    b = torch.as_tensor(b).float()
    ia = (a.view(torch.int32) >> 16).to(torch.int64)
    ib = (b.view(torch.int32) >> 16).to(torch.int64)
    return (ia - ib).abs()


def _scores(seed, S, m_block, gain=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(1, S, 1, 128, generator=g) * gain).to(torch.bfloat16)
    k = (torch.randn(1, S, 1, 128, generator=g) * gain).to(torch.bfloat16)
    v = torch.randn(1, S, 1, 128, generator=g).to(torch.bfloat16)
    # key tiles of three strengths in a fixed visit-order pattern: the running max climbs in steps, and weak tiles that follow a
    # strong one fall far below it (skip votes), the others do not (do votes)
    qt, kt = H.tiles(S)
    pattern = torch.tensor([1.0, 1.7, 0.3, 1.0, 1.7, 0.3, 0.3])
    ramp = pattern[(kt - 1 - torch.arange(kt)) % len(pattern)].repeat_interleave(BN)[:S]     # indexed by visit position
    k = (k.float() * ramp[None, :, None, None]).to(torch.bfloat16)
    qf = torch.zeros(qt * BM, 128)
    qf[:S] = q[0, :, 0].float()
    kf = torch.zeros(kt * BN, 128)
    kf[:S] = k[0, :, 0].float()
    Qt = qf[m_block * BM:(m_block + 1) * BM]
    order = list(range(kt - 1, -1, -1))
    S_tiles = torch.stack([Qt @ kf[n * BN:(n + 1) * BN].T for n in order])          # fp32, exact products
    return q, k, v, S_tiles, order


@needs_ref
@pytest.mark.parametrize("S,m_block,thr", [(1152, 2, -1.0), (1152, 8, -3.0), (2000, 15, -0.5), (704, 0, -10.0)])
def test_oracle_softmax_matches_reference_code(S, m_block, thr):
    """m, alpha, vote, bf16 P, LSE, 1/l of the oracle == the reference's own softmax.h / mask.h on the same S tiles.
    Covers a ragged first K tile (seqlen mask), the zero-padded last Q tile (m_block 8 of S=1152 has 128 valid rows,
    m_block 15 of S=2000 has 80) and thresholds where votes are mixed."""
    scale = 128 ** -0.5
    c = float(torch.tensor(scale, dtype=torch.float32) * torch.tensor(oa.LOG2E, dtype=torch.float32))
    q, k, v, S_tiles, order = _scores(11 + m_block, S, m_block)
    valid_first = S - order[0] * BN
    ref = run_reference_softmax(S_tiles.numpy(), order[0], S, S, m_block, c, thr)
    orc = oa.softmax_tile_sequence(list(S_tiles), c, scale, thr, first_tile_valid_cols=valid_first if valid_first < BN else None)
    n_votes = n_skip = 0
    for t, (r, o) in enumerate(zip(ref["tiles"], orc["tiles"])):
        assert np.array_equal(r["m_run"], o["m_run"].numpy()), f"tile {t}: running row max differs"      # max is exact
        # alpha: one ex2.approx (2 ulp) vs torch.exp2
        np.testing.assert_allclose(r["alpha"], o["alpha"].numpy(), rtol=4e-7 * 4, atol=1e-37, err_msg=f"tile {t}: alpha")
        if t > 0:
            tie = abs(o["stat"] - thr) < 1e-4 * max(1.0, abs(thr))
            assert tie or r["vote"] == o["vote"], f"tile {t}: vote differs (stat {o['stat']}, thr {thr})"
            n_votes += 1
            n_skip += int(r["vote"])
        # P: identical up to the reference's ex2.approx (<= 2 ulp of fp32) moving a value across a bf16 rounding
        # boundary: never more than one bf16 step, and rarely
        steps = bf16_ulp_steps(r["p"], o["p_bf16"].float())
        assert int(steps.max()) <= 1, f"tile {t}: bf16 P differs by more than one rounding step"
        assert float((steps > 0).float().mean()) < 2e-3, f"tile {t}: too many P elements differ"
    padded_q_tile = (m_block + 1) * BM > S      # zero-padded rows give stat 0 > thr: such a tile never votes skip
    assert 0 < n_skip < n_votes or thr <= -10.0 or (padded_q_tile and n_skip == 0), "test data should exercise both vote outcomes"
    np.testing.assert_allclose(ref["lse"], orc["lse"].numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(ref["inv"], orc["inv"].numpy(), rtol=2e-6)


@needs_ref
def test_kernel_statistic_votes_like_reference_code(native_lib):
    """The statistic la_fwd_kernel emits, thresholded, reproduces the vote of the reference's own
    max_get_scale_detect_qk_skip on every visited tile of several Q tiles (ties excepted)."""
    S, thr = 2000, -1.0
    scale = 128 ** -0.5
    c = float(torch.tensor(scale, dtype=torch.float32) * torch.tensor(oa.LOG2E, dtype=torch.float32))
    qt, kt = H.tiles(S)
    disagreements = ties = total = 0
    for m_block in (0, 7, 15):
        q, k, v, S_tiles, order = _scores(5, S, m_block)
        ref = run_reference_softmax(S_tiles.numpy(), order[0], S, S, m_block, c, thr)
        qd, kd, vd = (t.to(DEV) for t in (q, k, v))
        out = torch.empty_like(qd)
        lse = torch.empty(1, 1, S, dtype=torch.float32, device=DEV)
        stat = torch.full((1, 1, qt, kt), float("nan"), device=DEV)
        native_lib.fwd(qd, kd, vd, out, lse, scale, H.init_list(1, 1, qt, kt).to(DEV), stat)
        torch.cuda.synchronize()
        st = stat[0, 0, m_block].cpu()
        for t, n in enumerate(order):
            if t == 0:
                assert st[n] == float("inf")
                continue
            total += 1
            kernel_vote = not bool(st[n] > thr)
            if kernel_vote != ref["tiles"][t]["vote"]:
                disagreements += 1
                ties += int(abs(float(st[n]) - thr) < 1e-4 * max(1.0, abs(thr)))
        # LSE of the valid rows against the reference code's finalize()
        rows = min(BM, S - m_block * BM)
        got = lse[0, 0, m_block * BM:m_block * BM + rows].cpu().numpy()
        np.testing.assert_allclose(got, ref["lse"][:rows], rtol=0, atol=1e-3)
    assert disagreements == ties, f"{disagreements - ties} of {total} votes differ away from the threshold"
