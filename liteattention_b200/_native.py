"""ctypes binding of the C ABI in include/liteattn_b200.h (libliteattn_b200.so, built in-tree by
liteattention_b200/csrc/build.py).  There is NO fallback: if the library is missing or a call fails this
module raises -- the product path never routes through the oracle or a CPU implementation."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LITEATTN_B200_LIB: developer override used by tools/ to time experimental builds of the same ABI.
LIB_PATH = os.environ.get("LITEATTN_B200_LIB") or os.path.join(_HERE, "libliteattn_b200.so")

BLOCK_M = 128
BLOCK_N = 176
HEAD_DIM = 128

_c_i64 = ctypes.c_int64
_c_i32 = ctypes.c_int32
_c_vp = ctypes.c_void_p


class FwdParams(ctypes.Structure):
    _fields_ = (
        [("q", _c_vp), ("k", _c_vp), ("v", _c_vp), ("out", _c_vp), ("lse", _c_vp)]
        + [(f"{t}_{s}_stride", _c_i64) for t in "qkvo" for s in ("batch", "row", "head")]
        + [(n, _c_i32) for n in ("b", "h", "h_k", "seqlen_q", "seqlen_k", "d")]
        + [("softmax_scale", ctypes.c_float), ("read_list", _c_vp), ("tile_stat", _c_vp), ("out_is_f32", _c_i32),
           ("n_out_peers", _c_i32), ("out_rows_per_peer", _c_i32), ("reserved_", _c_i32), ("out_peer", _c_vp * 8)]
    )


class UpdateParams(ctypes.Structure):
    _fields_ = [
        ("read_list", _c_vp), ("must_do_list", _c_vp), ("write_list", _c_vp), ("tile_stat", _c_vp),
        ("b", _c_i32), ("h", _c_i32), ("qtiles", _c_i32), ("ktiles", _c_i32),
        ("thr", ctypes.c_float), ("overflow_count", _c_vp),
    ]


class CombineParams(ctypes.Structure):
    _fields_ = [
        ("o_parts", ctypes.POINTER(_c_vp)), ("lse_parts", ctypes.POINTER(_c_vp)), ("n_parts", _c_i32),
        ("out", _c_vp), ("lse", _c_vp),
        ("b", _c_i32), ("h", _c_i32), ("s", _c_i32), ("d", _c_i32),
        ("parts_are_f32", _c_i32), ("out_is_f32", _c_i32),
    ]


class RopeParams(ctypes.Structure):
    _fields_ = [
        ("x", _c_vp), ("out", _c_vp), ("cos_sin", _c_vp), ("grid", _c_vp),
        ("x_batch_stride", _c_i64), ("x_row_stride", _c_i64), ("x_head_stride", _c_i64),
        ("b", _c_i32), ("s", _c_i32), ("h", _c_i32), ("d", _c_i32), ("max_pos", _c_i32), ("x_is_bf16", _c_i32),
    ]


_lib = None

EXPORTS = ("la_abi_version", "la_last_error", "la_get_tile_mn", "la_fwd_sm100", "la_skip_update_sm100",
           "la_fwd_skip_sm100", "la_combine_sm100", "la_rope_cast_sm100", "la_launch_count", "la_list_pack_sm100",
           "la_list_unpack_sm100", "la_copy2d_async")


def lib():
    """Load (once) and return the native library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python liteattention_b200/csrc/build.py` "
                "(or __graft_entry__.build()). liteattention_b200 has no non-CUDA fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.la_abi_version.restype = ctypes.c_int
        L.la_last_error.restype = ctypes.c_char_p
        L.la_launch_count.restype = ctypes.c_uint64
        L.la_get_tile_mn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                     ctypes.POINTER(ctypes.c_int)]
        L.la_fwd_sm100.argtypes = [ctypes.POINTER(FwdParams), _c_vp]
        L.la_skip_update_sm100.argtypes = [ctypes.POINTER(UpdateParams), _c_vp]
        L.la_fwd_skip_sm100.argtypes = [ctypes.POINTER(FwdParams), ctypes.POINTER(UpdateParams), _c_vp]
        L.la_combine_sm100.argtypes = [ctypes.POINTER(CombineParams), _c_vp]
        L.la_rope_cast_sm100.argtypes = [ctypes.POINTER(RopeParams), _c_vp]
        L.la_watchdog_read.argtypes = [ctypes.POINTER(ctypes.c_uint * 4)]
        if hasattr(L, "la_list_pack_sm100"):
            L.la_list_pack_sm100.argtypes = [_c_vp, _c_vp, _c_i64, ctypes.c_int, _c_vp, _c_vp]
            L.la_list_unpack_sm100.argtypes = [_c_vp, _c_vp, _c_i64, ctypes.c_int, _c_vp]
        if hasattr(L, "la_copy2d_async"):
            L.la_copy2d_async.argtypes = [_c_vp, ctypes.c_size_t, _c_vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _c_vp]
        # LITEATTN_B200_LIB builds of older revisions (tools/ab.py) share the forward / update structs
        if L.la_abi_version() != 3 and not os.environ.get("LITEATTN_B200_LIB"):
            raise RuntimeError("libliteattn_b200.so ABI version mismatch")
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        msg = lib().la_last_error().decode()
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def get_tile_mn(head_dim, element_size=2, v_colmajor=False):
    m, n = ctypes.c_int(), ctypes.c_int()
    rc = lib().la_get_tile_mn(head_dim, element_size, int(v_colmajor), ctypes.byref(m), ctypes.byref(n))
    return m.value, n.value, rc == 0


class PeerScatter:
    """Sequence-parallel destination of O: query row r goes to peers[r // rows_per_peer][:, r % rows_per_peer].
    peers: list of (b, rows_per_peer, heads_here_or_more, d) bf16 views with identical strides (peer memory)."""

    def __init__(self, peers, rows_per_peer):
        assert 1 <= len(peers) <= 8 and all(t.dtype == torch.bfloat16 and t.stride() == peers[0].stride() for t in peers)
        self.peers, self.rows_per_peer = list(peers), int(rows_per_peer)
        self.dtype = torch.bfloat16

    def stride(self, i):
        return self.peers[0].stride(i)

    def data_ptr(self):
        return self.peers[0].data_ptr()


def make_fwd_params(q, k, v, out, lse, softmax_scale, read_list, tile_stat):
    p = FwdParams()
    if isinstance(out, PeerScatter):
        p.n_out_peers, p.out_rows_per_peer = len(out.peers), out.rows_per_peer
        for i, t in enumerate(out.peers):
            p.out_peer[i] = t.data_ptr()
    p.q, p.k, p.v, p.out, p.lse = _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse)
    for name, t in (("q", q), ("k", k), ("v", v), ("o", out)):
        setattr(p, f"{name}_batch_stride", t.stride(0))
        setattr(p, f"{name}_row_stride", t.stride(1))
        setattr(p, f"{name}_head_stride", t.stride(2))
    p.b, p.seqlen_q, p.h, p.d = q.shape
    p.seqlen_k, p.h_k = k.shape[1], k.shape[2]
    p.softmax_scale = float(softmax_scale)
    p.read_list = _ptr(read_list)
    p.tile_stat = _ptr(tile_stat)
    p.out_is_f32 = int(out.dtype == torch.float32)
    return p


def make_update_params(read_list, must_do_list, write_list, tile_stat, b, h, qtiles, ktiles, thr, overflow_count=None):
    u = UpdateParams()
    u.read_list, u.must_do_list, u.write_list = _ptr(read_list), _ptr(must_do_list), _ptr(write_list)
    u.tile_stat = _ptr(tile_stat)
    u.b, u.h, u.qtiles, u.ktiles = b, h, qtiles, ktiles
    u.thr = float(thr)
    u.overflow_count = _ptr(overflow_count)
    return u


def fwd(q, k, v, out, lse, softmax_scale, read_list=None, tile_stat=None):
    """la_fwd_sm100 on the current stream of q's device."""
    p = make_fwd_params(q, k, v, out, lse, softmax_scale, read_list, tile_stat)
    with torch.cuda.device(q.device):
        _check(lib().la_fwd_sm100(ctypes.byref(p), _stream(q.device)), "la_fwd_sm100")


def skip_update(read_list, must_do_list, write_list, tile_stat, b, h, qtiles, ktiles, thr, overflow_count=None):
    u = make_update_params(read_list, must_do_list, write_list, tile_stat, b, h, qtiles, ktiles, thr, overflow_count)
    with torch.cuda.device(tile_stat.device):
        _check(lib().la_skip_update_sm100(ctypes.byref(u), _stream(tile_stat.device)), "la_skip_update_sm100")


def fwd_skip(q, k, v, out, lse, softmax_scale, read_list, must_do_list, write_list, tile_stat, thr,
             overflow_count=None):
    """la_fwd_skip_sm100: forward gated by read_list, then write_list := update(read_list, stat, thr)."""
    b, sq, h, _ = q.shape
    qtiles = (sq + BLOCK_M - 1) // BLOCK_M
    ktiles = (k.shape[1] + BLOCK_N - 1) // BLOCK_N
    p = make_fwd_params(q, k, v, out, lse, softmax_scale, read_list, tile_stat)
    u = make_update_params(read_list, must_do_list, write_list, tile_stat, b, h, qtiles, ktiles, thr, overflow_count)
    with torch.cuda.device(q.device):
        _check(lib().la_fwd_skip_sm100(ctypes.byref(p), ctypes.byref(u), _stream(q.device)), "la_fwd_skip_sm100")


def combine(o_parts, lse_parts, out, lse):
    n = len(o_parts)
    b, s, h, d = out.shape
    c = CombineParams()
    oa = (_c_vp * n)(*[t.data_ptr() for t in o_parts])
    la = (_c_vp * n)(*[t.data_ptr() for t in lse_parts])
    c.o_parts, c.lse_parts, c.n_parts = oa, la, n
    c.out, c.lse = _ptr(out), _ptr(lse)
    c.b, c.h, c.s, c.d = b, h, s, d
    c.parts_are_f32 = int(o_parts[0].dtype == torch.float32)
    c.out_is_f32 = int(out.dtype == torch.float32)
    with torch.cuda.device(out.device):
        _check(lib().la_combine_sm100(ctypes.byref(c), _stream(out.device)), "la_combine_sm100")


def list_pack(lists, bits, bad_rows=None):
    """la_list_pack_sm100: int32 rows [..., ktiles+1] (contiguous) -> uint32 bits [rows, 2, ceil(ktiles/32)] (as int32 storage)."""
    ktiles = lists.shape[-1] - 1
    rows = lists.numel() // (ktiles + 1)
    with torch.cuda.device(lists.device):
        _check(lib().la_list_pack_sm100(_ptr(lists), _ptr(bits), rows, ktiles, _ptr(bad_rows), _stream(lists.device)),
               "la_list_pack_sm100")


def list_unpack(bits, lists):
    """la_list_unpack_sm100: bits [rows, 2, words] -> int32 rows [..., ktiles+1] (entries past len untouched)."""
    ktiles = lists.shape[-1] - 1
    rows = lists.numel() // (ktiles + 1)
    with torch.cuda.device(lists.device):
        _check(lib().la_list_unpack_sm100(_ptr(bits), _ptr(lists), rows, ktiles, _stream(lists.device)),
               "la_list_unpack_sm100")


def copy2d_async(dst, src, col0, ncols, stream):
    """la_copy2d_async for one column block of two equally shaped contiguous (rows..., cols) tensors, one of them in
    pinned host memory: elements [..., col0 : col0 + ncols] of `src` -> the same elements of `dst`, on `stream`."""
    assert dst.shape == src.shape and dst.dtype == src.dtype and dst.is_contiguous() and src.is_contiguous()
    cols = dst.shape[-1]
    rows = dst.numel() // cols
    es = dst.element_size()
    dev = dst.device if dst.is_cuda else src.device
    with torch.cuda.device(dev):
        _check(lib().la_copy2d_async(dst.data_ptr() + col0 * es, cols * es, src.data_ptr() + col0 * es, cols * es,
                                     ncols * es, rows, stream.cuda_stream), "la_copy2d_async")


def launch_count():
    return int(lib().la_launch_count())


def watchdog_read():
    arr = (ctypes.c_uint * 4)()
    lib().la_watchdog_read(ctypes.byref(arr))
    return list(arr)


def rope_cast(x, out, cos_sin, grid):
    """la_rope_cast_sm100: out (bf16) = rope3d(x) for x fp32 or bf16 (b, s, h, d); cos_sin fp32 [max_pos, d/2, 2];
    grid int32 [b, 3] on the device."""
    p = RopeParams()
    p.x, p.out, p.cos_sin, p.grid = _ptr(x), _ptr(out), _ptr(cos_sin), _ptr(grid)
    p.x_batch_stride, p.x_row_stride, p.x_head_stride = x.stride(0), x.stride(1), x.stride(2)
    p.b, p.s, p.h, p.d = x.shape
    p.max_pos = cos_sin.shape[0]
    p.x_is_bf16 = int(x.dtype == torch.bfloat16)
    with torch.cuda.device(x.device):
        _check(lib().la_rope_cast_sm100(ctypes.byref(p), _stream(x.device)), "la_rope_cast_sm100")
