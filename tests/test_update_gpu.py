"""GPU: the skip-list update kernel against the codec oracles, bit-exact (integer work + one fp32 compare)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from oracle import skiplist as sl
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "skiplist_ref")


def _random_case(rows, kt, seed, with_md, p_nan=0.02):
    rng = np.random.default_rng(seed)
    read = np.zeros((rows, kt + 1), np.int32)
    md = np.zeros((rows, kt + 1), np.int32)
    md[:, 0] = 2
    stat = rng.normal(size=(rows, kt)).astype(np.float32) * 3
    stat[rng.random((rows, kt)) < p_nan] = np.nan
    stat[rng.random((rows, kt)) < 0.02] = -np.inf
    assert kt >= 2          # a one-tile row cannot even hold one range
    for r in range(rows):
        while True:
            run = int(rng.integers(1, 6))
            keep = np.repeat(rng.random((kt + run - 1) // run) < rng.uniform(0.3, 1.0), run)[:kt]
            keep[kt - 1] = True
            row = sl.encode_keep_mask(keep.tolist())
            if len(row) <= kt + 1:
                break
        read[r, :len(row)] = row
        if with_md and r % 2 == 0:
            npts = 4 if kt >= 4 else 2
            pts = sorted(rng.integers(0, kt, npts).tolist(), reverse=True)
            md[r, :npts + 1] = [npts] + pts
    return read, md, stat


def _run_kernel(native_lib, read, md, stat, thr, b=1, h=1):
    rows, kp1 = read.shape
    kt = kp1 - 1
    rl = torch.from_numpy(read).to(DEV)
    mdl = None if md is None else torch.from_numpy(md).to(DEV)
    wl = torch.full_like(rl, -5)
    st = torch.from_numpy(stat).to(DEV)
    ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
    native_lib.skip_update(rl, mdl, wl, st, 1, 1, rows, kt, thr, ovf)
    torch.cuda.synchronize()
    return wl.cpu().numpy(), int(ovf.item())


@pytest.mark.parametrize("kt", [2, 7, 33, 187, 430, 1024, 1025, 2048])       # 1025 / 2048: more than one 32-word group
@pytest.mark.parametrize("with_md", [False, True])
def test_update_kernel_bit_exact_vs_c_oracle(native_lib, kt, with_md):
    rows = 3000 if kt < 100 else 1200 if kt < 1000 else 250
    read, md, stat = _random_case(rows, kt, seed=kt * 2 + with_md, with_md=with_md)
    for thr in (-1.0, 0.0, float("inf"), float("-inf")):
        got, n_ovf = _run_kernel(native_lib, read, md if with_md else None, stat, thr)
        exp, n_exp = H.c_oracle_step(read, md if with_md else None, stat, thr, on_overflow=1)
        assert H.rows_equal_upto_len(got, exp), (kt, with_md, thr)
        assert n_ovf == n_exp


def test_fast_and_general_paths_agree(native_lib):
    """A trivial must-do list ([2,0,0] rows) takes the ballot fast path, a must-do list that protects nothing
    ([2,k,k]) forces the serial path: same answer."""
    kt = 187
    read, md, stat = _random_case(2000, kt, seed=99, with_md=False)
    fast, _ = _run_kernel(native_lib, read, None, stat, -0.5)
    md2 = md.copy()
    md2[:, 1] = md2[:, 2] = 3                 # n <= 3 && n > 3 never holds
    slow, _ = _run_kernel(native_lib, read, md2, stat, -0.5)
    assert H.rows_equal_upto_len(fast, slow)


def test_unsorted_hand_made_list_and_overflow(native_lib):
    kt = 6
    # the overflow example of SURVEY Appendix A: read [4,5,3,2,0], skip votes {4,1}
    read = np.zeros((2, kt + 1), np.int32)
    read[0, :5] = [4, 5, 3, 2, 0]
    read[1, :5] = [4, 2, 0, 5, 3]             # ascending range order: legal for the reader, not for the fast path
    stat = np.full((2, kt), 1.0, np.float32)
    stat[0, [4, 1]] = -1.0
    stat[1, [1]] = -1.0
    got, n_ovf = _run_kernel(native_lib, read, None, stat, 0.0)
    assert n_ovf == 1 and got[0, :5].tolist() == [4, 5, 3, 2, 0]          # overflow policy: copy of the read row
    exp, _ = sl.skip_list_step(read[1].tolist(), lambda n: not (stat[1, n] > 0.0), None, kt)
    assert got[1, :len(exp)].tolist() == exp


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/skiplist_ref not built (needs /root/reference)")
@pytest.mark.parametrize("with_md", [False, True])
def test_codec_against_the_reference_structs(native_lib, with_md):
    """oracle/_ref/skiplist_ref runs the REFERENCE's own SkipListReader/SkipListWriter (compiled from
    /root/reference in the build container).  Its output pins (i) the C codec oracle and (ii) the CUDA update kernel."""
    kt, rows, thr = 61, 4000, 0.25
    read, md, stat = _random_case(rows, kt, seed=7 + with_md, with_md=with_md, p_nan=0.0)
    votes = (~(stat > thr)).astype(np.int32)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            np.array([rows, kt, int(with_md)], np.int32).tofile(f)
            read.tofile(f)
            md.tofile(f)
            votes.tofile(f)
        subprocess.check_call([REF_BIN, fin, fout], timeout=120)
        ref = np.fromfile(fout, np.int32).reshape(rows, 2 * (kt + 1))
    exp_unb, n_ovf = H.c_oracle_step(read, md if with_md else None, stat, thr, on_overflow=0)
    assert H.rows_equal_upto_len(ref, exp_unb), "C codec oracle disagrees with the reference's own structs"
    got, _ = _run_kernel(native_lib, read, md if with_md else None, stat, thr)
    fits = ref[:, 0] <= kt
    assert H.rows_equal_upto_len(got[fits], ref[fits][:, :kt + 1])
    assert (~fits).sum() == n_ovf


def test_single_tile_rows(native_lib):
    """ktiles == 1 (seqlen <= 176): rows are [len, 0]; the only tile is always visited and kept."""
    read = np.array([[2, 0], [0, 0]], np.int32)
    stat = np.array([[np.inf], [np.nan]], np.float32)
    got, _ = _run_kernel(native_lib, read, None, stat, 0.0)
    assert got[0].tolist() == [2, 0] and got[1, 0] == 0
    exp, _ = H.c_oracle_step(read, None, stat, 0.0)
    assert exp[0].tolist() == [2, 0] and exp[1, 0] == 0


@pytest.mark.gpu
def test_threshold_calibration_hits_a_target_sparsity(native_lib):
    """liteattention_b200.calibrate: bisection over the update kernel on a stored statistic; sparsity is monotone in
    thr and the calibrated list is what a LiteAttention object with that threshold writes."""
    import math
    from liteattention_b200 import LiteAttention
    from liteattention_b200.calibrate import calibrate_threshold
    b, s, h = 1, 2600, 2
    g = torch.Generator().manual_seed(5)
    t = torch.arange(s).float()
    w = torch.randn(128, generator=g) * 0.02
    ph = torch.rand(128, generator=g) * 2 * math.pi
    e = math.sqrt(2 / 128) * torch.cos(t[:, None] * w[None] + ph[None])
    base = 16.0 * e[None, :, None, :].expand(b, s, h, 128)
    q = (base + 0.5 * torch.randn(b, s, h, 128, generator=g)).to(torch.bfloat16).cuda()
    k = (base + 0.5 * torch.randn(b, s, h, 128, generator=g)).to(torch.bfloat16).cuda()
    v = torch.randn(b, s, h, 128, generator=g).to(torch.bfloat16).cuda()
    prev = -1.0
    thr_max, sp_max, _ = calibrate_threshold(q, k, v, 0.99)       # unreachable: the result at the upper bound comes back
    assert thr_max == -1e-3 and 0.3 < sp_max < 0.99
    for target in (0.1, 0.2, 0.3):
        thr, sp, wl = calibrate_threshold(q, k, v, target)
        assert thr < 0 and sp >= target and sp - target < 0.08, (target, thr, sp)
        assert sp >= prev
        prev = sp
        la = LiteAttention(enable_skipping=True, threshold=thr, max_batch_size=b)
        la(q, k, v)
        kt1 = wl.shape[-1]   # entries beyond row[0] are stale by design (the writer never clears them)
        assert H.rows_equal_upto_len(la.read_list[:b].cpu().view(-1, kt1).numpy(), wl.cpu().view(-1, kt1).numpy()), \
            "calibrated list != list written by LiteAttention at that threshold"
