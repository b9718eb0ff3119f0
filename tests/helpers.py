"""Shared helpers for the parity tests (test-side only; may import oracle/)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

from oracle import skiplist as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BM, BN = 128, 176


def tiles(sq, sk=None):
    sk = sq if sk is None else sk
    return (sq + BM - 1) // BM, (sk + BN - 1) // BN


def random_keep_list(b, h, qtiles, ktiles, p_keep, seed, force_last=True):
    """Fixed random Skip-Mask at tile granularity -> (read_list int32 [b,h,qtiles,ktiles+1], keep bool)."""
    g = torch.Generator().manual_seed(seed)
    keep = torch.rand(b, h, qtiles, ktiles, generator=g) < p_keep
    if force_last:
        keep[..., ktiles - 1] = True
    rl = torch.zeros(b, h, qtiles, ktiles + 1, dtype=torch.int32)
    kn = keep.numpy()
    for bi in range(b):
        for hi in range(h):
            for m in range(qtiles):
                row = sl.encode_keep_mask(kn[bi, hi, m].tolist())
                while len(row) > ktiles + 1:       # alternates too finely for the row format: close one gap
                    gap = next(n for n in range(ktiles) if not kn[bi, hi, m, n])
                    kn[bi, hi, m, gap] = True
                    row = sl.encode_keep_mask(kn[bi, hi, m].tolist())
                rl[bi, hi, m, :len(row)] = torch.tensor(row, dtype=torch.int32)
    return rl, keep


def init_list(b, h, qtiles, ktiles):
    rl = torch.zeros(b, h, qtiles, ktiles + 1, dtype=torch.int32)
    rl[..., 0] = 2
    rl[..., 1] = ktiles - 1
    return rl


_c_oracle = None


def c_oracle():
    """oracle/_build/libskiplist_oracle.so (built on demand with gcc)."""
    global _c_oracle
    if _c_oracle is None:
        path = os.path.join(ROOT, "oracle", "_build", "libskiplist_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        L.skiplist_oracle_step.restype = ctypes.c_int
        L.skiplist_oracle_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int]
        _c_oracle = L
    return _c_oracle


def c_oracle_step(read, must_do, stat, thr, on_overflow=1):
    """read/must_do int32 [rows, ktiles+1] (must_do may be None), stat fp32 [rows, ktiles] (numpy or CPU torch).
    Returns (written int32 [rows, stride], n_overflowed); stride = ktiles+1 (policy copy) or 2*(ktiles+1)."""
    read = np.ascontiguousarray(np.asarray(read, dtype=np.int32))
    stat = np.ascontiguousarray(np.asarray(stat, dtype=np.float32))
    rows, kp1 = read.shape
    ktiles = kp1 - 1
    stride = kp1 if on_overflow else 2 * kp1
    out = np.zeros((rows, stride), dtype=np.int32)
    mdp = None
    if must_do is not None:
        must_do = np.ascontiguousarray(np.asarray(must_do, dtype=np.int32))
        mdp = must_do.ctypes.data
    n = c_oracle().skiplist_oracle_step(read.ctypes.data, mdp, out.ctypes.data, stat.ctypes.data, rows, ktiles,
                                        float(thr), stride, int(on_overflow))
    return out, n


def rows_equal_upto_len(a, b):
    """Compare list rows on [0, len] only (entries past len are stale by design, SkipListWriter :184-191)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape[0] != b.shape[0]:
        return False
    for ra, rb in zip(a, b):
        if ra[0] != rb[0]:
            return False
        n = int(ra[0])
        if not np.array_equal(ra[1:1 + n], rb[1:1 + n]):
            return False
    return True


def fa_tolerance(out, out_ref_fp32, out_ref_lowp):
    """FlashAttention's tolerance idiom (hopper/tests/test_flash_attn.py:266-296 of the reference):
    |out - ref| <= 2 * |ref_bf16 - ref| + fwd_atol with fwd_atol = 2 * |(ref + 0.3 - 0.3) - ref|max."""
    fwd_atol = 2 * (out_ref_fp32 + 0.3 - 0.3 - out_ref_fp32).abs().max().item()
    err = (out.float() - out_ref_fp32).abs().max().item()
    err_lowp = (out_ref_lowp.float() - out_ref_fp32).abs().max().item()
    return err, 2 * err_lowp + fwd_atol
