"""CPU: pins the oracle against the reference-generated fixtures, the hand-simulated codec vectors and the
reference's own smoke invariants (test_lite_attention.py:12-92 of the reference)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import attention as oa
from oracle import skiplist as sl
from tests import helpers as H


@pytest.fixture(scope="module")
def book(golden_dir):
    return json.load(open(os.path.join(golden_dir, "host_bookkeeping.json")))


@pytest.fixture(scope="module")
def codec(golden_dir):
    return json.load(open(os.path.join(golden_dir, "codec_vectors.json")))


def test_get_mn_matches_reference(book):
    for e in book["get_MN"]:
        assert list(sl.get_MN(e["head_dim"], e["element_size"], e["v_colmajor"])) == e["mn"]


def test_init_row_matches_reference(book):
    for e in book["init_skip_list"]:
        b, s, h, d = e["args"]
        bm, bn = sl.get_MN(d, 2)
        qt, kt = sl.ceil_div(s, bm), sl.ceil_div(s, bn)
        assert e["shape"] == [2, b, h, qt, kt + 1]
        assert sl.init_row(kt)[:4] == e["row_prefix"]
        assert e["all_rows_equal"]


def test_expand_must_do_matches_reference(book, codec):
    for e in book["expand_must_do"]:
        kt = sl.ceil_div(e["seq_len"], 176)
        row = sl.expand_must_do(e["must_do_list"], kt)
        assert len(row) == kt + 1 == e["shape"][3]
        assert row[:len(e["row_prefix"])] == e["row_prefix"]
    ex = codec["expand_must_do_example"]
    assert sl.expand_must_do(ex["must_do_list"], 12, ex["k_tile_size"])[:7] == ex["row"]


def _vote_fn(votes):
    if votes == "all":
        return lambda n: True
    s = set(votes)
    return lambda n: n in s


def test_codec_vectors(codec):
    for v in codec["vectors"]:
        kt = v["ktiles"]
        rd = v["read"] + [0] * (kt + 1 - len(v["read"]))
        md = v["must_do"] + [0] * (kt + 1 - len(v["must_do"]))
        row, vis = sl.skip_list_step(rd, _vote_fn(v["skip_votes"]), md, kt, on_overflow="unbounded")
        assert row == v["written"]
        assert vis == v["visited"]
        if v.get("overflows"):
            with pytest.raises(sl.Overflow):
                sl.skip_list_step(rd, _vote_fn(v["skip_votes"]), md, kt, on_overflow="raise")
            row2, _ = sl.skip_list_step(rd, _vote_fn(v["skip_votes"]), md, kt, on_overflow="copy")
            assert row2 == v["read"]


def test_c_oracle_matches_python_oracle():
    rng = np.random.default_rng(0)
    for kt in (6, 7, 12, 33, 100):
        rows = 200
        read = np.zeros((rows, kt + 1), np.int32)
        md = np.zeros((rows, kt + 1), np.int32)
        stat = rng.normal(size=(rows, kt)).astype(np.float32)
        stat[rng.random((rows, kt)) < 0.05] = np.nan
        for r in range(rows):
            while True:
                keep = rng.random(kt) < rng.uniform(0.2, 1.0)
                keep[kt - 1] = True
                row = sl.encode_keep_mask(keep.tolist())
                if len(row) <= kt + 1:  # a finely alternating mask does not fit the row format
                    break
            read[r, :len(row)] = row
            if r % 3 == 0:
                a, b2 = sorted(rng.integers(0, kt, 2).tolist(), reverse=True)
                md[r, :3] = [2, a, b2]
            else:
                md[r, :3] = [2, 0, 0]
        thr = 0.0
        for on_overflow, mode in ((1, "copy"), (0, "unbounded")):
            got, _ = H.c_oracle_step(read, md, stat, thr, on_overflow)
            for r in range(rows):
                exp, _ = sl.skip_list_step(read[r].tolist(), lambda n, r=r: not (stat[r, n] > thr), md[r].tolist(),
                                           kt, on_overflow=mode)
                assert got[r, :len(exp)].tolist() == exp, (kt, r, mode)


def test_codec_structural_properties():
    """written lists have even length, non-increasing entries, and the tiles listed at
    step t+1 are a subset of those visited at step t (sparsity is monotone)."""
    rng = np.random.default_rng(1)
    kt = 9
    for _ in range(300):
        rd = sl.init_row(kt)
        prev_vis = None
        for step in range(4):
            votes = {n: bool(rng.random() < 0.5) for n in range(kt)}
            try:
                row, vis = sl.skip_list_step(rd, lambda n: votes[n], None, kt, on_overflow="raise")
            except sl.Overflow:
                break
            assert row[0] % 2 == 0
            ent = row[1:1 + row[0]]
            assert all(ent[i] >= ent[i + 1] for i in range(len(ent) - 1))
            nxt = set(sl.visited_tiles(row + [0] * (kt + 1 - len(row)), kt))
            assert nxt <= set(vis)
            if prev_vis is not None:
                assert set(vis) <= prev_vis
            prev_vis = set(vis)
            rd = row + [0] * (kt + 1 - len(row))


# ---- the reference's own smoke invariants (test_lite_attention.py), on the oracle ------------------------------
@pytest.fixture(scope="module")
def small_qkv():
    torch.manual_seed(0)
    b, s, h, d = 1, 700, 2, 128
    return tuple(torch.randn(b, s, h, d).to(torch.bfloat16) for _ in range(3))


def test_invariant_skip_all(small_qkv):
    q, k, v = small_qkv
    qt, kt = H.tiles(q.shape[1])
    r = oa.lite_attention_oracle(q, k, v, None, H.init_list(1, 2, qt, kt), None, thr=float("inf"))
    wl = r["write_list"]
    assert (wl[..., 0] <= 2).all()                                  # :12-27
    assert (wl[..., 1] == kt - 1).all() and (wl[..., 2] == kt - 2).all()


def test_invariant_must_do_everything(small_qkv):
    q, k, v = small_qkv
    s = q.shape[1]
    qt, kt = H.tiles(s)
    rl = H.init_list(1, 2, qt, kt)
    md = torch.tensor(sl.expand_must_do([s - 1, 0], kt), dtype=torch.int32).expand(1, 2, qt, kt + 1).contiguous()
    r = oa.lite_attention_oracle(q, k, v, None, rl, md, thr=float("inf"))
    assert torch.equal(r["write_list"], rl)                         # :29-45


def test_invariant_skip_nothing(small_qkv):
    q, k, v = small_qkv
    qt, kt = H.tiles(q.shape[1])
    rl = H.init_list(1, 2, qt, kt)
    r = oa.lite_attention_oracle(q, k, v, None, rl, None, thr=float("-inf"))
    assert torch.equal(r["write_list"], rl)                         # :47-56


def test_invariant_lse_and_dense_output(small_qkv):
    q, k, v = small_qkv
    qt, kt = H.tiles(q.shape[1])
    r = oa.lite_attention_oracle(q, k, v, None, H.init_list(1, 2, qt, kt), None, thr=0.0)
    o_ref, lse_ref = oa.dense_attention_ref(q, k, v)
    assert (r["lse"] - lse_ref).abs().max() < 1e-3                  # reference bound is 0.1 (:58-92)
    err, tol = H.fa_tolerance(r["out"], o_ref, o_ref.to(torch.bfloat16))
    assert err <= tol


def test_config1_dense_vs_sdpa():
    """BASELINE config 1: dense S=2048 d=64 fp32 1 head vs torch SDPA on CPU (plumbing check of the fp32 reference
    used throughout the parity tests)."""
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, 2048, 1, 64) for _ in range(3))
    o, _ = oa.dense_attention_ref(q, k, v)
    ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    assert (o - ref.transpose(1, 2)).abs().max() < 1e-5
