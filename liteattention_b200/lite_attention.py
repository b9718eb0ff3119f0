"""Stateful wrapper around the B200 QK-Skip attention forward.

Host-side mirror of hopper/lite_attention.py of the reference (class LiteAttention :15-320,
SeqParallelLiteAttention :322-345): same constructor, call signature, public attributes (`threshold`,
`enable_skipping`, `_skip_list` [2, max_batch, H, qtiles, ktiles+1] int32, `_phase`), same re-initialisation
rules and the same list format, so that code written against the reference (README.md:265-323, the Wan
integration) runs unchanged.  Differences, all documented in DESIGN.md:
  * `enable_skipping=False` runs the dense kernel (the reference intends that, README.md:159, but raises on
    `None.shape` because lite_attention.py:262 tests a bound method);
  * the default (empty) must-do list is not materialised per call (the reference allocates and broadcasts a
    [max_batch,H,qtiles,ktiles+1] tensor every call, :239-241); the kernel treats "no list" as `[2,0,0]`;
  * `calc_percentage` keeps the reference's (broken for descending lists, :61-85) formula for drop-in
    compatibility; `sparsity()` / `last_sparsity` give the correct figure;
  * `compact_state=True` (keyword-only extension, or LITE_ATTENTION_COMPACT_STATE=1; off by default): the resident
    state is two bits per (row, K tile), allocated for the batch actually seen -- 2.6 MB per layer object at the
    Wan2.1-14B shape instead of the reference's 326 MB int32 double buffer (:124, max_batch_size 4).  The int32 rows
    the kernels consume live in a scratch pair shared by all layer objects of a stream; la_list_unpack / la_list_pack
    (csrc/la_list_codec.cu) convert around every call, losslessly.
  * host-resident activations (extension): called with PINNED CPU tensors, the object streams them through the GPU by
    head groups -- the forward of group g runs while group g+1 is on the wire and group g-1's O is on its way back
    (la_copy2d_async) -- and returns O in pinned host memory.  There is still no CPU compute path: pageable CPU
    tensors raise.
"""
import os
from typing import Optional, Tuple, Union

import torch

from . import _native
from .flash_attn_interface import _flash_attn_forward, flash_attn_func, fwd_peer_scatter


_LIST_WS = {}


def _list_scratch(device, shape):
    """(read, write) int32 scratch rows shared by every compact-state LiteAttention object of a device / stream / shape:
    a call expands its bitmaps into `read`, the kernels write `write`, the object packs `write` back -- all on one
    stream, so the next object can reuse the pair."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tuple(shape))
    ws = _LIST_WS.get(key)
    if ws is None:
        if len(_LIST_WS) >= 8:
            _LIST_WS.clear()
        ws = _LIST_WS[key] = (torch.zeros(shape, dtype=torch.int32, device=device),
                              torch.zeros(shape, dtype=torch.int32, device=device))
    return ws


class _HostStage:
    """Device staging for calls on pinned host tensors: two slots of (q, k, v, o) device buffers and of pinned host
    outputs (call n uses slot n & 1, so its uploads overlap call n-1's compute), one upload and one download stream.
    Shared by every layer object of a device / shape / dtype."""

    def __init__(self, device, shape, dtype):
        self.dev = [[torch.empty(shape, dtype=dtype, device=device) for _ in range(4)] for _ in range(2)]
        self.host_out = [None, None]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]      # slot reusable: its last download has finished
        self.h2d = torch.cuda.Stream(device)
        self.d2h = torch.cuda.Stream(device)
        self.calls = 0
        self.last = None                                           # completion event of the most recent call


_HOST_STAGES = {}


def _host_stage(device, shape, dtype):
    key = (device.index, tuple(shape), dtype)
    st = _HOST_STAGES.get(key)
    if st is None:
        if len(_HOST_STAGES) >= 2:
            # The staging buffers are used by copies on side streams the caching allocator knows nothing about: drain
            # everything before their memory can be handed out again.
            for old in _HOST_STAGES.values():
                old.h2d.synchronize()
                old.d2h.synchronize()
            torch.cuda.synchronize(device)
            _HOST_STAGES.clear()
        st = _HOST_STAGES[key] = _HostStage(device, shape, dtype)
    return st


def host_head_groups(heads: int, busy: bool = False):
    """Head-group sizes for streamed host calls: small first and last groups (the first upload and the last download
    are the only copies nothing overlaps), growing by 2x in between so that a group's upload always finishes under
    the previous group's compute.  `busy`: the previous host call is still in flight -- its compute already covers
    this call's uploads, so one big group + the small last one (for a short drain) is enough and saves launches.
    LITE_ATTENTION_HOST_CHUNKS="2,4,8,..." overrides (must sum to `heads`)."""
    env = os.getenv("LITE_ATTENTION_HOST_CHUNKS", "")
    if env:
        sizes = [int(x) for x in env.split(",")]
        if sum(sizes) != heads or min(sizes) <= 0:
            raise ValueError(f"LITE_ATTENTION_HOST_CHUNKS={env!r} does not partition {heads} heads")
        return sizes
    if heads <= 3:
        return [heads]
    last = max(1, heads // 10)
    if busy:
        return [heads - last, last]
    sizes, left, nxt = [], heads - last, max(1, heads // 20)
    while left > 0:
        g = min(nxt, left, max(1, heads // 3))
        sizes.append(g)
        left -= g
        nxt *= 2
    return sizes + [last]


class LiteAttention:
    """Attention with temporal-sparse QK-Skip lists managed internally (one object per layer, reused across
    diffusion timesteps).

    Args:
        enable_skipping: gate K tiles by the skip list and keep updating it.  Default True.
        threshold: QK-skip threshold in the exp2 domain (a tile is dropped once every row's largest softmax
            numerator relative to the running max is <= 2**threshold).  Must be negative unless
            LITE_ATTENTION_DEBUG is set.  Default -10.0.
        max_batch_size: leading dimension the list buffers are allocated for.  Default 4.
    """

    def __init__(self, enable_skipping: bool = True, threshold: float = -10.0, max_batch_size: int = 4, *,
                 compact_state: Optional[bool] = None):
        self._compact = (os.getenv("LITE_ATTENTION_COMPACT_STATE", "0") not in ("0", "", "FALSE", "false")
                         if compact_state is None else bool(compact_state))
        self._bits = None          # compact mode: int32-typed storage of uint32 [batch, H, qtiles, 2, words]
        self._skip_list_buf = None
        self._phase = 0

        self._last_seq_len = None
        self._last_head_dim = None
        self._last_v_colmajor = None
        self._last_dtype = None
        self._last_device = None
        self._last_num_heads = None

        self._last_percentage = 0.0
        self._must_do_cache_key = None
        self._must_do_cache = None

        self.enable_skipping = enable_skipping
        self.set_threshold(threshold)
        self.max_batch_size = max_batch_size

    # ------------------------------------------------------------------ state
    @property
    def _skip_list(self):
        """[2, batch, H, qtiles, ktiles+1] int32, like the reference's attribute.  In compact mode it is exported on
        demand from the bitmaps and both halves hold the list the next call will read."""
        if self._compact:
            if self._bits is None:
                return None
            cur = self._export_lists()
            return torch.stack([cur, cur])
        return self._skip_list_buf

    @_skip_list.setter
    def _skip_list(self, value):
        if self._compact and value is not None:
            lists = value[self._phase] if value.dim() == 5 else value
            self._import_lists(lists.contiguous())
        elif self._compact:
            self._bits = None
        else:
            self._skip_list_buf = value

    def _export_lists(self, batch: Optional[int] = None):
        b = self._bits.shape[0] if batch is None else batch
        h, qt, kt = self._bits.shape[1], self._bits.shape[2], self._ktiles
        out = torch.zeros(b, h, qt, kt + 1, dtype=torch.int32, device=self._bits.device)
        _native.list_unpack(self._bits[:b], out)
        return out

    def _import_lists(self, lists: torch.Tensor):
        """Pack int32 rows [batch, H, qtiles, ktiles+1] into the resident bitmaps (rows must be descending/disjoint)."""
        b, h, qt, kp1 = lists.shape
        words = (kp1 - 1 + 31) // 32
        bits = torch.empty(b, h, qt, 2, words, dtype=torch.int32, device=lists.device)
        bad = torch.zeros(1, dtype=torch.int32, device=lists.device)
        _native.list_pack(lists, bits, bad)
        if int(bad.item()) != 0:
            raise ValueError("compact_state: the skip list has rows that are not descending, disjoint ranges; "
                             "use compact_state=False for hand-made lists")
        self._bits, self._ktiles, self._qtiles = bits, kp1 - 1, qt

    # ------------------------------------------------------------------ static helpers (reference API)
    @staticmethod
    def ceil_div(x, y):
        return (x + y - 1) // y

    @staticmethod
    def calc_percentage(read_list: torch.Tensor) -> float:
        """Reference formula (hopper/lite_attention.py:61-85), kept verbatim in meaning: it was written for
        ascending lists and returns a NEGATIVE "not skipped" fraction for the descending lists the kernel
        uses.  Prefer `LiteAttention.sparsity`."""
        read_list = read_list.to(torch.int64)
        skip_lengths = read_list[:, :, :, 0] // 2
        sized = read_list[:, :, :, 2:] - read_list[:, :, :, 1:-1]
        if sized.shape[-1] % 2 != 0:
            sized = torch.cat([sized, torch.zeros_like(sized[..., :1])], dim=-1)
        sized = sized.view(*sized.shape[:3], -1, 2)[..., 0].cumsum(dim=-1)
        total_possible = read_list.shape[0] * read_list.shape[1] * read_list.shape[2] * (read_list.shape[3] - 1)
        # gather index len//2 is one past the last range; the reference does the same (clamped here for safety)
        idx = skip_lengths.clamp(max=sized.shape[-1] - 1).unsqueeze(-1)
        total_not_skipped = torch.gather(sized, dim=-1, index=idx).squeeze(-1).sum()
        return total_not_skipped / total_possible if total_possible > 0 else 1.0

    @staticmethod
    def sparsity(skip_list: torch.Tensor) -> float:
        """Fraction of (q-tile, k-tile) pairs NOT listed: 1 - sum_ranges(start - end + 1) / ktiles, averaged over
        rows.  skip_list: [..., ktiles+1] int32 rows of [len, s0, e0, ...]."""
        rows = skip_list.reshape(-1, skip_list.shape[-1]).to(torch.int64)
        ktiles = rows.shape[1] - 1
        ln = rows[:, 0].clamp(0, ktiles)
        ent = rows[:, 1:]
        pos = torch.arange(ktiles, device=rows.device)
        valid = pos[None, :] < ln[:, None]
        sign = torch.where(pos % 2 == 0, 1, -1)[None, :]           # +start, -end
        covered = (ent * sign * valid).sum(dim=1) + ln // 2
        return float(1.0 - covered.double().mean().item() / ktiles)

    @staticmethod
    def get_MN(head_dim, element_size, v_colmajor=False):
        """(kBlockM, kBlockN) of the skip-list tiles; same table as the reference (hopper/lite_attention.py:87-111,
        tile_size.h:10-62).  The list geometry is API, so the B200 kernel gates on these Hopper tile sizes."""
        if element_size == 2:
            if head_dim <= 64:
                return 192, 192
            elif head_dim <= 96:
                return 192, 144
            elif head_dim <= 128:
                return 128, 176
            elif head_dim <= 192:
                return 128, 112
            else:
                return 128, 80
        else:
            if head_dim <= 64:
                return 192, 160
            elif head_dim <= 96:
                return 192, 128
            elif head_dim <= 128:
                return 128, (192 if v_colmajor else 224)
            elif head_dim <= 192:
                return 128, 160
            else:
                return 128, 128

    @staticmethod
    def init_skip_list(batch, seq_len, heads, head_dim, v_colmajor, dtype, device, must_skip_list=None) -> torch.Tensor:
        """[2, batch, heads, qtiles, ktiles+1] int32 double buffer, every row = one range over all K tiles
        (hopper/lite_attention.py:113-153).  `must_skip_list` (token ranges that are never computed) is the
        reference's experimental branch (:126-145); it is reproduced without mutating the caller's list."""
        element_size = dtype.itemsize
        kTileM, kTileN = LiteAttention.get_MN(head_dim, element_size, v_colmajor)
        qtiles = LiteAttention.ceil_div(seq_len, kTileM)
        ktiles = LiteAttention.ceil_div(seq_len, kTileN)
        skip_list = torch.zeros(2, batch, heads, qtiles, ktiles + 1, dtype=torch.int32, device=device)
        if must_skip_list is not None:
            msl = list(must_skip_list)
            n = msl[0] if msl else 0          # the reference reads element 0 as a length here (:130)
            for i in range(1, min(n, len(msl) - 1) + 1):
                if i % 2 == 1:
                    msl[i] = (msl[i] + kTileN - 1) // kTileN
                else:
                    msl[i] = msl[i] // kTileN
            msl.insert(0, ktiles - 1)
            msl.append(0)
            msl.insert(0, len(msl))
            values = torch.tensor(msl, dtype=torch.int32, device=device)
            skip_list[:, :, :, :, :len(msl)] = values
        else:
            skip_list[:, :, :, :, 1] = ktiles - 1
            skip_list[:, :, :, :, 0] = 2
        return skip_list

    def _init_skip_list(self, query, value, must_skip_list=None):
        batch, seq_len, heads, head_dim = query.shape
        assert batch <= self.max_batch_size, \
            "batch size must be less than or equal to max_batch_size (modify max_batch_size in LiteAttention constructor)"
        v_colmajor = value.shape[-3] == head_dim
        return LiteAttention.init_skip_list(self.max_batch_size, seq_len, heads, head_dim, v_colmajor,
                                            query.dtype, query.device, must_skip_list)

    def _get_read_write_lists(self, query, value, must_skip_list=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Pick this call's (read, write) halves of the double buffer; (re)initialise on any shape / dtype /
        device change; flip the phase (hopper/lite_attention.py:165-212)."""
        if not self.enable_skipping:
            return None, None
        current_seq_len = query.shape[1]
        head_dim = query.shape[-1]
        current_num_heads = query.shape[2]
        v_colmajor = value.shape[-3] == head_dim
        dtype, device = query.dtype, query.device
        state = self._bits if self._compact else self._skip_list_buf
        if (state is None or self._last_seq_len != current_seq_len
                or state.device != query.device or self._last_head_dim != head_dim
                or self._last_v_colmajor != v_colmajor or self._last_dtype != dtype
                or self._last_device != device or self._last_num_heads != current_num_heads):
            if self._compact:
                batch, seq_len, heads, _ = query.shape
                assert batch <= self.max_batch_size, \
                    "batch size must be less than or equal to max_batch_size (modify max_batch_size in LiteAttention constructor)"
                # allocated for the batch actually seen (grown on demand below), not for max_batch_size
                self._import_lists(LiteAttention.init_skip_list(batch, seq_len, heads, head_dim, v_colmajor, dtype, device,
                                                                must_skip_list)[0])
            else:
                self._skip_list_buf = self._init_skip_list(query, value, must_skip_list)
            self._phase = 0
            self._last_seq_len = current_seq_len
            self._last_head_dim = head_dim
            self._last_v_colmajor = v_colmajor
            self._last_dtype = dtype
            self._last_device = device
            self._last_num_heads = current_num_heads
            if os.getenv("LITE_ATTENTION_VERBOSE", "FALSE") != "FALSE":
                print("[Warning]: reinitialized skip list during the forward pass")
        if self._compact:
            batch = query.shape[0]
            assert batch <= self.max_batch_size, \
                "batch size must be less than or equal to max_batch_size (modify max_batch_size in LiteAttention constructor)"
            if batch > self._bits.shape[0]:        # new batch rows start from the dense initial list
                extra = LiteAttention.init_skip_list(batch - self._bits.shape[0], current_seq_len, current_num_heads, head_dim,
                                                     v_colmajor, dtype, device)[0]
                old = self._bits
                self._import_lists(extra)
                self._bits = torch.cat([old, self._bits])
            read_list, write_list = _list_scratch(query.device, (batch,) + tuple(self._bits.shape[1:3]) + (self._ktiles + 1,))
            _native.list_unpack(self._bits[:batch], read_list)
            self._phase ^= 1
            return read_list, write_list
        if self._phase == 0:
            read_list, write_list = self._skip_list_buf[0], self._skip_list_buf[1]
            self._phase = 1
        else:
            read_list, write_list = self._skip_list_buf[1], self._skip_list_buf[0]
            self._phase = 0
        return read_list, write_list

    @staticmethod
    def _expand_must_do_list(must_do_list, list_shape, query, value):
        """1-D token ranges [s0, e0, s1, e1, ...] -> int32 [batch, heads, qtiles, ktiles+1] block-range rows
        `[len, ceil(s/kN), floor(e/kN), ...]` (hopper/lite_attention.py:214-242; the rounding quirk is kept)."""
        head_dim = query.shape[-1]
        v_colmajor = value.shape[-3] == head_dim
        q_tile_size, k_tile_size = LiteAttention.get_MN(head_dim, query.dtype.itemsize, v_colmajor)
        must_do_list = [len(must_do_list)] + list(must_do_list)
        for i in range(1, must_do_list[0] + 1):
            if i % 2 == 1:
                must_do_list[i] = (must_do_list[i] + k_tile_size - 1) // k_tile_size
            else:
                must_do_list[i] = must_do_list[i] // k_tile_size
        values = torch.tensor(must_do_list, dtype=torch.int32, device=query.device)
        values = torch.cat([values, torch.zeros(list_shape[3] - values.size(0), dtype=values.dtype, device=values.device)])
        return values.repeat(*list_shape[:3], 1).contiguous()

    def _must_do_expanded(self, must_do_list, list_shape, query, value):
        key = (tuple(must_do_list), tuple(list_shape), query.device, query.dtype, query.shape[-1])
        if self._must_do_cache_key != key:
            self._must_do_cache = self._expand_must_do_list(list(must_do_list), list_shape, query, value)
            self._must_do_cache_key = key
        return self._must_do_cache

    # ------------------------------------------------------------------ call
    def __call__(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, scale: Optional[float] = None,
                 return_softmax_lse: bool = False, must_do_list: list = None,
                 must_skip_list: list = None, *, out: Optional[torch.Tensor] = None
                 ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        """query/key/value: (batch, seq_len, heads, head_dim) bf16.  Returns the attention output
        (batch, seq_len, heads, head_dim) [and softmax_lse (batch, heads, seq_len) fp32].
        must_do_list: token ranges [start0, end0, start1, end1, ...] (descending) that may never be skipped.
        out (extension, keyword-only; the reference's op takes it, `flash_api.cpp:861-870`, its Python does not pass
        it): write the result there instead of allocating -- any device-visible bf16 (batch, seq_len, heads,
        head_dim) buffer, including a peer GPU's memory mapped over NVLink (liteattention_b200/dist.py)."""
        if query.device.type == "cpu" and query.is_pinned():
            # host-resident activations: streamed through the GPU by head groups (pageable CPU tensors fall through to
            # the op, which has no CPU kernel and raises)
            return self._call_host(query, key, value, scale, return_softmax_lse, must_do_list, must_skip_list, out)
        read_list, write_list = self._get_read_write_lists(query, value, must_skip_list)

        must_do_list_expanded = None
        if self.enable_skipping and must_do_list is not None:
            must_do_list_expanded = self._must_do_expanded(must_do_list, write_list.shape, query, value)

        if isinstance(out, _native.PeerScatter):
            # Sequence-parallel return path (liteattention_b200/dist.py): O rows are scattered to the peers that own
            # their tokens by the kernel's epilogue.  Not expressible through the reference's op (its `out` is one
            # tensor), so this goes to the C ABI directly with the same list handling.
            lse = fwd_peer_scatter(query, key, value, out, scale, read_list, must_do_list_expanded, write_list,
                                   self.threshold)
            output = (out, lse) if return_softmax_lse else out
        elif out is None:
            output = flash_attn_func(q=query, k=key, v=value, softmax_scale=scale, attn_read_list=read_list,
                                     attn_must_do_list=must_do_list_expanded, attn_write_list=write_list,
                                     thr=self.threshold, return_softmax_lse=return_softmax_lse)
        else:
            o, lse, *_ = _flash_attn_forward(
                query, key, value, None, None, None, out, None, None, None, None, None, None, None, None, None, None,
                None, None, None, None, None, None,
                scale if scale is not None else query.shape[-1] ** (-0.5), causal=False,
                attn_read_list=read_list, attn_must_do_list=must_do_list_expanded, attn_write_list=write_list,
                thr=self.threshold)
            output = (o, lse) if return_softmax_lse else o

        if self._compact and write_list is not None:
            _native.list_pack(write_list, self._bits[:query.shape[0]])      # the list the next call will read
        if self.enable_skipping and os.getenv("LITE_ATTENTION_VERBOSE", "FALSE") != "FALSE":
            real_batch_size = query.shape[0]
            self._last_percentage = 1.0 - LiteAttention.sparsity(read_list[:real_batch_size])
            print(f"[Info]: Percentage of tiles skipped: {1.0 - self._last_percentage:.2%}")
        return output

    # ------------------------------------------------------------------ call on pinned host tensors
    def _call_host(self, query, key, value, scale, return_softmax_lse, must_do_list, must_skip_list, out):
        """query/key/value in PINNED host memory, contiguous (batch, seq_len, heads, head_dim) bf16.  The tensors are
        streamed through the current CUDA device by head groups; the list handling, kernels and results are those of
        a device call (same bits).  Returns O in pinned host memory (`out` if given, else one of two internal buffers
        that alternate between calls) [and softmax_lse on the device].  Everything is asynchronous: call
        `join_host_copies()` (stream-ordered) or `wait_host_copies()` (blocking) before reading the result on the CPU
        or overwriting the inputs."""
        for name, t in (("query", query), ("key", key), ("value", value)):
            if t.device.type != "cpu" or not t.is_pinned():
                raise RuntimeError(f"LiteAttention: {name} is a pageable CPU tensor; host-resident inputs must be pinned "
                                   "(liteattention_b200 computes on the GPU only, there is no CPU path)")
            if not t.is_contiguous():
                raise ValueError(f"LiteAttention: host-resident {name} must be contiguous (batch, seq_len, heads, head_dim)")
        if not (query.shape == key.shape == value.shape) or query.dim() != 4:
            raise ValueError("LiteAttention: host-resident q, k, v must share one (batch, seq_len, heads, head_dim) shape")
        if isinstance(out, _native.PeerScatter):
            raise NotImplementedError("LiteAttention: PeerScatter outputs need device-resident inputs")
        if not torch.cuda.is_available():
            raise RuntimeError("LiteAttention: no CUDA device (liteattention_b200 has no CPU path)")
        device = torch.device("cuda", torch.cuda.current_device())
        B, S, H, D = query.shape
        st = _host_stage(device, query.shape, query.dtype)
        slot = st.calls & 1
        st.calls += 1
        dq, dk, dv, do = st.dev[slot]
        if out is None:
            if st.host_out[slot] is None:
                st.host_out[slot] = torch.empty(query.shape, dtype=query.dtype).pin_memory()
            out = st.host_out[slot]
        elif (out.device.type != "cpu" or not out.is_pinned() or not out.is_contiguous() or out.shape != query.shape
              or out.dtype != query.dtype):
            raise ValueError("LiteAttention: `out` for host-resident inputs must be a pinned contiguous CPU tensor of q's shape")

        read_list, write_list = self._get_read_write_lists(dq, dv, must_skip_list)
        md = None
        if self.enable_skipping and must_do_list is not None:
            md = self._must_do_expanded(must_do_list, write_list.shape, dq, dv)
        softmax_scale = scale if scale is not None else D ** (-0.5)
        lse = torch.empty((B, H, S), dtype=torch.float32, device=device) if return_softmax_lse else None

        main = torch.cuda.current_stream(device)
        st.h2d.wait_event(st.free[slot])          # the call that used this slot two calls ago has drained
        cols = H * D
        flat = lambda t: t.view(B * S, cols)
        busy = st.last is not None and not st.last.query()     # some layer object's host call is still in flight
        h0 = 0
        for g in host_head_groups(H, busy):
            h1 = h0 + g
            for dst, src in ((dq, query), (dk, key), (dv, value)):
                _native.copy2d_async(flat(dst), flat(src), h0 * D, g * D, st.h2d)
            ev_in = torch.cuda.Event()
            ev_in.record(st.h2d)
            main.wait_event(ev_in)
            for b in range(B):                    # list slices of one batch row and a head range are contiguous
                sl = (slice(b, b + 1), slice(h0, h1))
                _, lse_g, *_ = _flash_attn_forward(
                    dq[b:b + 1, :, h0:h1], dk[b:b + 1, :, h0:h1], dv[b:b + 1, :, h0:h1], None, None, None,
                    do[b:b + 1, :, h0:h1], None, None, None, None, None, None, None, None, None, None,
                    None, None, None, None, None, None, softmax_scale, causal=False,
                    attn_read_list=None if read_list is None else read_list[sl],
                    attn_must_do_list=None if md is None else md[sl],
                    attn_write_list=None if write_list is None else write_list[sl], thr=self.threshold)
                if lse is not None:
                    lse[b:b + 1, h0:h1].copy_(lse_g)
            ev_out = torch.cuda.Event()
            ev_out.record(main)
            st.d2h.wait_event(ev_out)
            _native.copy2d_async(flat(out), flat(do), h0 * D, g * D, st.d2h)
            h0 = h1
        st.free[slot].record(st.d2h)
        self._host_done = st.last = st.free[slot]
        if self._compact and write_list is not None:
            _native.list_pack(write_list, self._bits[:B])
        return (out, lse) if return_softmax_lse else out

    def join_host_copies(self):
        """Make the current stream wait for the downloads of the last host-resident call (stream-ordered)."""
        ev = getattr(self, "_host_done", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def wait_host_copies(self):
        """Block until the result of the last host-resident call is in host memory."""
        ev = getattr(self, "_host_done", None)
        if ev is not None:
            ev.synchronize()

    @property
    def read_list(self) -> Optional[torch.Tensor]:
        """The list the NEXT call will read (i.e. the one the last call wrote)."""
        if self._compact:
            return None if self._bits is None else self._export_lists()
        if self._skip_list_buf is None:
            return None
        return self._skip_list_buf[self._phase]

    def last_sparsity(self, batch: Optional[int] = None) -> float:
        """Sparsity of the list the next call will use."""
        if (self._bits if self._compact else self._skip_list_buf) is None:
            return 0.0
        rl = self.read_list
        return LiteAttention.sparsity(rl if batch is None else rl[:batch])

    def load_skip_list(self, skip_list: torch.Tensor, query: torch.Tensor, value: torch.Tensor):
        """Extension (not in the reference): start from a given list instead of the dense initial one, e.g. a list
        exported from an earlier run or synthesised at a target sparsity.  skip_list: int32
        [batch <= max_batch_size, heads, qtiles, ktiles+1] for the geometry of `query`."""
        if self._compact:
            self._import_lists(skip_list.to(device=query.device, dtype=torch.int32).contiguous())
        else:
            buf = self._init_skip_list(query, value)
            assert skip_list.shape[1:] == buf.shape[2:] and skip_list.shape[0] <= buf.shape[1], \
                f"skip_list shape {tuple(skip_list.shape)} does not match {tuple(buf.shape[1:])}"
            buf[:, :skip_list.shape[0]] = skip_list.to(device=buf.device, dtype=torch.int32)
            self._skip_list_buf = buf
        self._phase = 0
        self._last_seq_len = query.shape[1]
        self._last_head_dim = query.shape[-1]
        self._last_v_colmajor = value.shape[-3] == query.shape[-1]
        self._last_dtype = query.dtype
        self._last_device = query.device
        self._last_num_heads = query.shape[2]

    def reset_skip_state(self):
        """Forget the skip lists (next call starts dense).  hopper/lite_attention.py:293-304."""
        self._skip_list_buf = None
        self._bits = None
        self._phase = 0
        self._last_seq_len = None
        self._last_head_dim = None
        self._last_v_colmajor = None
        self._last_dtype = None
        self._last_device = None
        self.verbose_reinitialization = False
        self._last_percentage = 0.0
        self._last_num_heads = None

    def set_threshold(self, threshold: float):
        """Threshold must be negative unless LITE_ATTENTION_DEBUG is set (hopper/lite_attention.py:306-313)."""
        if threshold >= 0 and os.getenv("LITE_ATTENTION_DEBUG", "FALSE") == "FALSE":
            raise ValueError("threshold must be negative when debug mode is not enabled")
        self.threshold = threshold

    def enable_skip_optimization(self, enable: bool = True):
        self.enable_skipping = enable


class SeqParallelLiteAttention:
    """`num_nodes` independent LiteAttention states, one per K/V shard (`split_idx`), for callers that shard the
    sequence and merge partial results by LSE (hopper/lite_attention.py:322-345; merge: flash_attn_combine)."""

    def __init__(self, num_nodes: int, enable_skipping: bool = True, threshold: float = -10.0, max_batch_size: int = 4,
                 *, compact_state: Optional[bool] = None):
        self.num_nodes = num_nodes
        self.lite_attention = [LiteAttention(enable_skipping, threshold, max_batch_size, compact_state=compact_state)
                               for _ in range(num_nodes)]
        self.set_threshold(threshold)

    def __call__(self, query, key, value, split_idx: int, scale: Optional[float] = None,
                 return_softmax_lse: bool = False):
        assert split_idx < self.num_nodes, "split_idx must be less than num_nodes"
        return self.lite_attention[split_idx](query, key, value, scale, return_softmax_lse)

    def reset_skip_state(self):
        for la in self.lite_attention:
            la.reset_skip_state()

    def set_threshold(self, threshold: float):
        for la in self.lite_attention:
            la.set_threshold(threshold)

    def enable_skip_optimization(self, enable: bool = True):
        for la in self.lite_attention:
            la.enable_skip_optimization(enable)
