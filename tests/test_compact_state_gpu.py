"""GPU: compact resident skip state: two bits per (row, K tile) instead of the reference's int32
double buffer, lossless against the list codec oracle, and invisible in the results."""
import numpy as np
import pytest
import torch

from oracle import skiplist as sl
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rows_with_adjacent_ranges(kt, rows, seed):
    """Random descending, disjoint range lists, INCLUDING ranges that touch (e.g. [11..8][7..3]): the update kernel
    writes those and their boundaries matter (the writer's state is reset at every range start)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((rows, kt + 1), np.int32)
    for r in range(rows):
        n = kt - 1
        ent = []
        while n >= 0 and len(ent) + 2 <= kt:
            length = int(rng.integers(1, 6))
            s, e = n, max(n - length + 1, 0)
            ent += [s, e]
            gap = int(rng.integers(0, 4))            # 0 = the next range touches this one
            n = e - 1 - gap
            if rng.random() < 0.05:
                break
        out[r, 0] = len(ent)
        out[r, 1:1 + len(ent)] = ent
    return out


@pytest.mark.parametrize("kt", [1, 2, 7, 33, 64, 187, 430])
def test_pack_unpack_round_trip(native_lib, kt):
    rows = 2000
    lists = _rows_with_adjacent_ranges(kt, rows, seed=kt) if kt > 1 else np.array([[2, 0], [0, 0]] * 4, np.int32)
    rows = lists.shape[0]
    # the tiles each row visits, by the oracle's reader
    want_tiles = [sl.visited_tiles(lists[r].tolist(), kt) for r in range(rows)]
    dl = torch.from_numpy(lists).to(DEV)
    words = (kt + 31) // 32
    bits = torch.full((rows, 2, words), -1, dtype=torch.int32, device=DEV)
    bad = torch.zeros(1, dtype=torch.int32, device=DEV)
    native_lib.list_pack(dl, bits, bad)
    back = torch.full_like(dl, -7)
    native_lib.list_unpack(bits, back)
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
    assert H.rows_equal_upto_len(back.cpu().numpy(), lists)                       # lossless on [0, len]
    vis = bits[:, 0].cpu().numpy().view(np.uint32)
    for r in range(0, rows, 97):                                                  # the bitmap IS the visited set
        got = [n for n in range(kt - 1, -1, -1) if (vis[r, n >> 5] >> (n & 31)) & 1]
        assert got == want_tiles[r]


def test_pack_rejects_unsorted_rows(native_lib):
    kt = 12
    lists = torch.tensor([[4, 5, 3, 9, 7, 0, 0, 0, 0, 0, 0, 0, 0],      # ascending ranges: not representable
                          [2, 11, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], dtype=torch.int32, device=DEV)
    bits = torch.zeros(2, 2, 1, dtype=torch.int32, device=DEV)
    bad = torch.zeros(1, dtype=torch.int32, device=DEV)
    native_lib.list_pack(lists, bits, bad)
    assert int(bad.item()) == 1


def test_compact_trajectory_equals_default(native_lib):
    """12 chained calls with evolving QK-Skip: the compact-state object produces the same O and, after every step, the
    same list bits as the default int32 double buffer."""
    from liteattention_b200 import LiteAttention
    from tests.test_fwd_gpu import _local_qk
    b, s, h, thr = 2, 2600, 2, -10.0
    qt, kt = H.tiles(s)
    a = LiteAttention(threshold=thr, max_batch_size=4)
    c = LiteAttention(threshold=thr, max_batch_size=4, compact_state=True)
    for step in range(12):
        q, k, v = (t.to(DEV) for t in _local_qk(b, s, h, seed=3 + step))
        oa_, oc_ = a(q, k, v), c(q, k, v)
        assert torch.equal(oa_, oc_), step
        la, lc = a.read_list[:b].cpu().view(-1, kt + 1).numpy(), c.read_list.cpu().view(-1, kt + 1).numpy()
        assert H.rows_equal_upto_len(la, lc), step
    assert c.last_sparsity(b) > 0.15
    assert c._skip_list.shape == (2, b, h, qt, kt + 1)            # still inspectable in the reference's layout
    # batch growth: new rows start dense, old rows keep their lists
    q3, k3, v3 = (t.to(DEV) for t in _local_qk(3, s, h, seed=99))
    c(q3, k3, v3)
    assert c._bits.shape[0] == 3


def test_compact_state_size_at_the_wan_shape(native_lib):
    """Resident bytes per layer object at B=1, S=75600, H=40: <= 25 MB (the reference: 326 MB at max_batch_size 4)."""
    from liteattention_b200 import LiteAttention
    c = LiteAttention(compact_state=True)
    lists = LiteAttention.init_skip_list(1, 75600, 40, 128, False, torch.bfloat16, DEV)[0]
    c._import_lists(lists)
    resident = c._bits.numel() * c._bits.element_size()
    assert resident <= 25e6 and resident == 40 * 591 * 2 * 14 * 4
    default = 2 * 4 * 40 * 591 * 431 * 4
    assert default > 300e6
    assert torch.equal(c._export_lists(), lists)
