"""Pure-Python restatement of the reference's run-length skip-list codec and host bookkeeping.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

List row format (SkipListReader, hopper/_internal/cpp/mainloop_fwd_sm90_tma_gmma_ws.hpp:47-115):
    row = [len, s0, e0, s1, e1, ...]   len = number of entries; ranges inclusive, s >= e,
    tiles are visited s, s-1, ..., e, ranges in list order.
"""

BLOCK_M = 128
BLOCK_N = 176


def ceil_div(x, y):
    return (x + y - 1) // y


def get_MN(head_dim, element_size, v_colmajor=False):
    """tile_size_fwd_sm90 (hopper/_internal/cpp/tile_size.h:10-62) == LiteAttention.get_MN
    (hopper/lite_attention.py:87-111), non-causal / non-local branch."""
    if element_size == 2:
        table = ((64, (192, 192)), (96, (192, 144)), (128, (128, 176)), (192, (128, 112)))
        default = (128, 80)
    else:
        table = ((64, (192, 160)), (96, (192, 128)), (128, (128, 192 if v_colmajor else 224)), (192, (128, 160)))
        default = (128, 128)
    for lim, mn in table:
        if head_dim <= lim:
            return mn
    return default


def init_row(ktiles):
    """init_skip_list (hopper/lite_attention.py:113-153, the no-must-skip branch :147-151):
    one range covering every tile."""
    row = [0] * (ktiles + 1)
    row[0] = 2
    row[1] = ktiles - 1
    return row


def expand_must_do(must_do_list, ktiles, k_tile_size=BLOCK_N):
    """_expand_must_do_list (hopper/lite_attention.py:214-242) for one row: token ranges -> block ranges,
    odd positions (starts) rounded UP, even positions (ends) rounded DOWN, zero padded to ktiles+1."""
    lst = [len(must_do_list)] + list(must_do_list)
    for i in range(1, lst[0] + 1):
        if i % 2 == 1:
            lst[i] = (lst[i] + k_tile_size - 1) // k_tile_size
        else:
            lst[i] = lst[i] // k_tile_size
    return lst + [0] * (ktiles + 1 - len(lst))


def visited_tiles(read_row, ktiles=None):
    """Tiles the forward visits for this row, in order (mainloop :1804-1827).  With ktiles given, ranges are
    clamped the way the CUDA kernel clamps them (the reference does not validate list contents)."""
    ln = read_row[0]
    if ktiles == 1:          # one-tile rows are [len, 0]: no room for the range end, tile 0 is always visited
        return [0] if ln > 0 else []
    if ktiles is not None:
        ln = min(max(ln, 0), ktiles) & ~1
    out = []
    for r in range(0, ln, 2):
        s, e = read_row[1 + r], read_row[2 + r]
        if ktiles is not None:
            s, e = min(s, ktiles - 1), max(e, 0)
        out.extend(range(s, e - 1, -1))
    return out


class Overflow(Exception):
    """The reference writer would write past the end of the row."""


def skip_list_step(read_row, vote_skip, must_do_row=None, ktiles=None, on_overflow="raise"):
    """One forward call's effect on the skip list of one (b, h, q-tile) row.

    vote_skip(n) -> bool is the RAW tile vote `!any_row((m_loc - m_prev) * c > thr)` (softmax.h:194,207) for a
    visited tile n; it is never asked about the first visited tile (mainloop :1804-1805).
    Restates SkipListWriter (mainloop :121-192) and the range loop (:1804-1827), including:
      * must-do reader advanced by a single `if` (:156-159);
      * record_range_end receives the RAW vote (:1812-1816);
      * `skip` is declared once and not reset per range.
    Returns (written_row_entries [len, ...], visited tiles).  Capacity = ktiles entries (row has ktiles+1 ints);
    on_overflow: "raise" -> Overflow, "copy" -> the CUDA policy (written row := read row), "unbounded" -> keep.
    """
    if ktiles is None:
        ktiles = len(read_row) - 1
    if ktiles == 1:
        return ([2, 0], [0]) if read_row[0] > 0 else ([0], [])
    md = list(must_do_row) if must_do_row is not None else [2, 0, 0]

    def md_at(i):
        return md[i] if 0 <= i < len(md) and i <= ktiles else 0

    mdlen = md_at(0)
    mi, ms, me = 1, md_at(1), md_at(2)
    out = []
    skipping = True
    raw = False
    first = True
    visited = []
    ln = min(max(read_row[0], 0), ktiles) & ~1
    for r in range(0, ln, 2):
        s, e = min(read_row[1 + r], ktiles - 1), max(read_row[2 + r], 0)
        if s < e:
            continue
        for n in range(s, e - 1, -1):
            visited.append(n)
            if first:
                vote = raw = False
                first = False
            else:
                raw = bool(vote_skip(n))
                vote = raw
                if vote:
                    if me > n and mi <= mdlen:
                        mi += 2
                        ms, me = md_at(mi), md_at(mi + 1)
                    if n <= ms and n > me:
                        vote = False
            if vote != skipping:
                out.append(n)
                skipping = vote
        skipping = True
        if not raw:
            out.append(e)
    if len(out) > ktiles:
        if on_overflow == "raise":
            raise Overflow(f"{len(out)} entries for {ktiles} slots")
        if on_overflow == "copy":
            out = list(read_row[1:1 + ln])
    return [len(out)] + out, visited


def sparsity(read_row, ktiles):
    """Fraction of K tiles NOT visited (the reference's calc_percentage is wrong for descending lists,
    hopper/lite_attention.py:61-85)."""
    return 1.0 - len(set(visited_tiles(read_row, ktiles))) / ktiles


def encode_keep_mask(keep):
    """keep[n] (bool per K tile) -> list row entries [len, s0, e0, ...] with maximal descending runs."""
    ent = []
    n = len(keep) - 1
    while n >= 0:
        if keep[n]:
            s = n
            while n - 1 >= 0 and keep[n - 1]:
                n -= 1
            ent += [s, n]
        n -= 1
    return [len(ent)] + ent
