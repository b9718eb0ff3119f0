"""Functional API + PyTorch op boundary of the B200 QK-Skip attention forward.

Mirrors hopper/_internal/flash_attn_interface.py of the reference for this path:
  flash_attn_func        :547-635     FlashAttnFunc.forward :277-342     _flash_attn_forward :20-112
and registers the dispatcher op `lite_attention::fwd` with the schema of
hopper/_internal/cpp/flash_api.cpp:1722-1763 (TORCH_LIBRARY(lite_attention, m)).  The op's CUDA implementation
validates like mha_fwd (flash_api.cpp:667-1249, skip args :915-963) and then calls the C ABI
(include/liteattn_b200.h) on the current stream.  No CPU / eager fallback exists: on a non-CUDA tensor the
dispatcher raises, and a missing libliteattn_b200.so raises at first use.
"""
from typing import Optional

import torch

from . import _native

__all__ = ["flash_attn_func", "flash_attn_combine", "FlashAttnFunc", "_flash_attn_forward", "maybe_contiguous",
           "fwd_peer_scatter"]

_FWD_SCHEMA = (
    "fwd("
    "Tensor q,"
    "Tensor k,"
    "Tensor v,"
    "Tensor(k_new!)? k_new = None,"
    "Tensor(v_new!)? v_new = None,"
    "Tensor? q_v = None,"
    "Tensor(out!)? out = None,"
    "Tensor? cu_seqlens_q = None,"
    "Tensor? cu_seqlens_k = None,"
    "Tensor? cu_seqlens_k_new = None,"
    "Tensor? seqused_q = None,"
    "Tensor? seqused_k = None,"
    "int? max_seqlen_q = None,"
    "int? max_seqlen_k = None,"
    "Tensor? page_table = None,"
    "Tensor? kv_batch_idx = None,"
    "Tensor? leftpad_k = None,"
    "Tensor? rotary_cos = None,"
    "Tensor? rotary_sin = None,"
    "Tensor? seqlens_rotary = None,"
    "Tensor? q_descale = None,"
    "Tensor? k_descale = None,"
    "Tensor? v_descale = None,"
    "float? softmax_scale = None,"
    "bool is_causal = False,"
    "int window_size_left = -1,"
    "int window_size_right = -1,"
    "int attention_chunk = 0,"
    "float softcap = 0.0,"
    "bool is_rotary_interleaved = False,"
    "Tensor? scheduler_metadata = None,"
    "int num_splits = 0,"
    "bool? pack_gqa = None,"
    "int sm_margin = 0,"
    "Tensor? attn_read_list = None,"
    "Tensor? attn_must_do_list = None,"
    "Tensor? attn_write_list = None,"
    "float thr = -3.0) -> (Tensor(out!), Tensor, Tensor, Tensor)"
)

# Ops the reference registers next to fwd (flash_api.cpp:1764-1816), schema strings identical.  bwd throws in the
# reference's shipped build too (:1251-1254, DISABLE_BACKWARD); get_scheduler_metadata only serves the varlen /
# split schedulers that are compiled out.  fwd_combine is implemented (la_combine_sm100).
_BWD_SCHEMA = (
    "bwd("
    "Tensor dout,"
    "Tensor q,"
    "Tensor k,"
    "Tensor v,"
    "Tensor out,"
    "Tensor softmax_lse,"
    "Tensor(dq!)? dq = None,"
    "Tensor(dk!)? dk = None,"
    "Tensor(dv!)? dv = None,"
    "Tensor? cu_seqlens_q = None,"
    "Tensor? cu_seqlens_k = None,"
    "Tensor? seqused_q = None,"
    "Tensor? seqused_k = None,"
    "int? max_seqlen_q = None,"
    "int? max_seqlen_k = None,"
    "float? softmax_scale = None,"
    "bool is_causal = False,"
    "int window_size_left = -1,"
    "int window_size_right = -1,"
    "float softcap = 0.0,"
    "bool deterministic = False,"
    "int sm_margin = 0) -> (Tensor(dq!), Tensor(dk!), Tensor(dv!), Tensor, Tensor, Tensor, Tensor, Tensor)"
)
_COMBINE_SCHEMA = (
    "fwd_combine("
    "Tensor out_partial,"
    "Tensor lse_partial,"
    "Tensor(out!)? out = None,"
    "ScalarType? out_dtype = None) -> (Tensor(out!), Tensor)"
)
_SCHED_SCHEMA = (
    "get_scheduler_metadata("
    "int batch_size,"
    "int max_seqlen_q,"
    "int max_seqlen_k,"
    "int num_heads,"
    "int num_heads_k,"
    "int headdim,"
    "int headdim_v,"
    "ScalarType qkv_dtype,"
    "Tensor seqused_k,"
    "Tensor? cu_seqlens_q = None,"
    "Tensor? cu_seqlens_k = None,"
    "Tensor? cu_seqlens_k_new = None,"
    "Tensor? seqused_q = None,"
    "Tensor? leftpad_k = None,"
    "int? page_size = None,"
    "int max_seqlen_k_new = 0,"
    "bool is_causal = False,"
    "int window_size_left = -1,"
    "int window_size_right = -1,"
    "int attention_chunk = 0,"
    "bool has_softcap = False,"
    "int num_splits = 0,"
    "bool? pack_gqa = None,"
    "int sm_margin = 0) -> Tensor"
)
_STUB_SCHEMAS = {"bwd": _BWD_SCHEMA, "get_scheduler_metadata": _SCHED_SCHEMA}


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _check_list(t, name, b, h, qtiles, ktiles, device):
    # flash_api.cpp:919-963 checks dtype/dim/contiguity only; geometry checks are ours (a malformed list is an
    # out-of-bounds tile index in the reference).
    _check(t.dtype == torch.int32, f"{name} must be int32 tensor")
    _check(t.dim() == 4, f"{name} must be 4D tensor with shape [batch, heads, q_blocks, k_blocks]")
    _check(t.is_contiguous(), f"{name} must be contiguous")
    _check(t.device == device, f"{name} must be on the same device as q")
    _check(t.shape[0] >= b and t.shape[1] == h and t.shape[2] == qtiles and t.shape[3] == ktiles + 1,
           f"{name} has shape {tuple(t.shape)}, expected [>={b}, {h}, {qtiles}, {ktiles + 1}] "
           f"(tiles of {_native.BLOCK_M}x{_native.BLOCK_N})")


def _fwd_cuda(q, k, v, k_new=None, v_new=None, q_v=None, out=None, cu_seqlens_q=None, cu_seqlens_k=None,
              cu_seqlens_k_new=None, seqused_q=None, seqused_k=None, max_seqlen_q=None, max_seqlen_k=None,
              page_table=None, kv_batch_idx=None, leftpad_k=None, rotary_cos=None, rotary_sin=None,
              seqlens_rotary=None, q_descale=None, k_descale=None, v_descale=None, softmax_scale=None,
              is_causal=False, window_size_left=-1, window_size_right=-1, attention_chunk=0, softcap=0.0,
              is_rotary_interleaved=False, scheduler_metadata=None, num_splits=0, pack_gqa=None, sm_margin=0,
              attn_read_list=None, attn_must_do_list=None, attn_write_list=None, thr=-3.0):
    # -- features of the upstream op that the LiteAttention path never uses (most are compiled out of the
    #    reference's shipped build as well, hopper/setup.py:47-68, flash_api.cpp:1199-1216)
    for name, val in (("k_new", k_new), ("v_new", v_new), ("q_v", q_v), ("cu_seqlens_q", cu_seqlens_q),
                      ("cu_seqlens_k", cu_seqlens_k), ("cu_seqlens_k_new", cu_seqlens_k_new),
                      ("seqused_q", seqused_q), ("seqused_k", seqused_k), ("page_table", page_table),
                      ("kv_batch_idx", kv_batch_idx), ("leftpad_k", leftpad_k), ("rotary_cos", rotary_cos),
                      ("rotary_sin", rotary_sin), ("seqlens_rotary", seqlens_rotary), ("q_descale", q_descale),
                      ("k_descale", k_descale), ("v_descale", v_descale),
                      ("scheduler_metadata", scheduler_metadata)):
        if val is not None:
            raise NotImplementedError(f"lite_attention::fwd (sm_100a): argument '{name}' is not supported")
    if is_causal or window_size_left != -1 or window_size_right != -1 or attention_chunk != 0 or softcap != 0.0:
        raise NotImplementedError("lite_attention::fwd (sm_100a): causal/local/chunked/softcap attention is not "
                                  "supported (DiT self-attention is non-causal)")
    if num_splits > 1:
        raise NotImplementedError("lite_attention::fwd (sm_100a): num_splits > 1 is not supported")

    q, k, v, softmax_scale = _validate_qkv(q, k, v, softmax_scale)
    b, sq, h, d = q.shape
    if out is None:
        out = torch.empty((b, sq, h, d), dtype=q.dtype, device=q.device)                          # :871-875
    else:
        # extension: an fp32 `out` receives the bf16-rounded result widened (what the caller's `x.float()` would give)
        _check(out.dtype in (q.dtype, torch.float32) and out.shape == (b, sq, h, d) and out.stride(-1) == 1 and
               out.device == q.device and
               out.data_ptr() % 16 == 0 and all((s_ * out.element_size()) % 16 == 0 for s_ in out.stride()[:3]),
               "out must be (batch, seqlen_q, heads, head_dim) bf16 (or fp32) on q's device with contiguous, "
               "16-byte aligned rows")
    lse = _launch_fwd(q, k, v, out, softmax_scale, attn_read_list, attn_must_do_list, attn_write_list, thr)
    empty = q.new_empty(0)
    return out, lse, empty, empty.float()


def _validate_qkv(q, k, v, softmax_scale):
    """The tensor checks of mha_fwd (flash_api.cpp:700-856) for the one configuration built here; returns q, k, v made
    TMA-addressable (16-byte aligned base, strides multiples of 8 elements) and the default scale."""
    _check(q.is_cuda and k.is_cuda and v.is_cuda, "q, k, v must be CUDA tensors")
    _check(q.device == k.device and q.device == v.device, "q, k, v must be on the same device")
    _check(q.dtype == torch.bfloat16, "lite_attention::fwd (sm_100a) only supports bf16")        # setup.py:54-55
    _check(k.dtype == q.dtype and v.dtype == q.dtype, "query, key and value must have the same dtype")
    _check(q.dim() == 4 and k.dim() == 4 and v.dim() == 4, "q, k, v must be (batch, seqlen, heads, head_dim)")
    _check(q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1,
           "Input tensor must have contiguous last dimension")                                   # :726-728
    b, sq, h, d = q.shape
    sk, hk = k.shape[1], k.shape[2]
    _check(k.shape == (b, sk, hk, d) and v.shape == (b, sk, hk, d), "k/v shape mismatch")         # :807-826
    _check(h % hk == 0, "Number of heads in key/value must divide number of heads in query")
    _check(d % 8 == 0, "head_dim must be a multiple of 8")                                        # :854-856
    if d != _native.HEAD_DIM:
        raise NotImplementedError(f"lite_attention::fwd (sm_100a): head_dim {d} is not built (only 128)")
    if softmax_scale is None:
        softmax_scale = d ** -0.5

    def _tma_ok(t):
        return t.data_ptr() % 16 == 0 and all(s % 8 == 0 for s in t.stride()[:3])
    q, k, v = (t if _tma_ok(t) else t.contiguous() for t in (q, k, v))
    return q, k, v, softmax_scale


# Scratch for the per-tile statistic (forward writes it, the update kernel reads it, same stream, same call): one
# buffer per (device, stream, shape) shared by every layer object instead of a 40 MB torch.empty per call.
_STAT_WS = {}


def _stat_workspace(device, shape):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tuple(shape))
    ws = _STAT_WS.get(key)
    if ws is None:
        if len(_STAT_WS) >= 16:
            _STAT_WS.clear()
        ws = _STAT_WS[key] = torch.empty(shape, dtype=torch.float32, device=device)
    return ws


def _launch_fwd(q, k, v, out, softmax_scale, attn_read_list, attn_must_do_list, attn_write_list, thr):
    """List validation (flash_api.cpp:915-963 + geometry) and the C-ABI call; `out` is a tensor or a
    _native.PeerScatter.  q, k, v must have been through _validate_qkv.  Returns softmax_lse."""
    b, sq, h, d = q.shape
    sk = k.shape[1]
    lse = torch.empty((b, h, sq), dtype=torch.float32, device=q.device)                           # :887-892
    qtiles = (sq + _native.BLOCK_M - 1) // _native.BLOCK_M
    ktiles = (sk + _native.BLOCK_N - 1) // _native.BLOCK_N
    if attn_read_list is not None:                       # is_skipable (flash_api.cpp:919-936)
        _check_list(attn_read_list, "attn_read_list", b, h, qtiles, ktiles, q.device)
        if attn_must_do_list is not None:
            _check_list(attn_must_do_list, "attn_must_do_list", b, h, qtiles, ktiles, q.device)
        if attn_write_list is not None:
            _check_list(attn_write_list, "attn_write_list", b, h, qtiles, ktiles, q.device)
            _check(attn_write_list.data_ptr() != attn_read_list.data_ptr(),
                   "attn_read_list and attn_write_list must not alias")
            stat = _stat_workspace(q.device, (b, h, qtiles, ktiles))
            _native.fwd_skip(q, k, v, out, lse, softmax_scale, attn_read_list, attn_must_do_list,
                             attn_write_list, stat, thr)
        else:
            _native.fwd(q, k, v, out, lse, softmax_scale, attn_read_list, None)
    else:
        _native.fwd(q, k, v, out, lse, softmax_scale, None, None)
    return lse


def fwd_peer_scatter(q, k, v, scatter, softmax_scale=None, attn_read_list=None, attn_must_do_list=None,
                     attn_write_list=None, thr=-3.0):
    """The forward with O scattered by rows to peer GPUs (liteattention_b200/dist.py).  Not expressible through the
    reference's op (its `out` is one tensor); same validation as the op, then the C ABI.  Returns softmax_lse."""
    q, k, v, softmax_scale = _validate_qkv(q, k, v, softmax_scale)
    b, sq, h, d = q.shape
    _check(isinstance(scatter, _native.PeerScatter), "scatter must be a PeerScatter")
    p0 = scatter.peers[0]
    _check(all(t.device.type == "cuda" and t.dim() == 4 and t.shape[0] >= b and t.shape[2] == h and t.shape[3] == d
               and t.stride(-1) == 1 and t.shape[1] >= min(scatter.rows_per_peer, sq) for t in scatter.peers)
           and scatter.rows_per_peer * len(scatter.peers) >= sq and p0.data_ptr() % 16 == 0
           and all(s_ % 8 == 0 for s_ in p0.stride()[:3]),
           "PeerScatter buffers must be (batch, rows_per_peer, heads, head_dim) bf16 views covering seqlen_q")
    return _launch_fwd(q, k, v, scatter, softmax_scale, attn_read_list, attn_must_do_list, attn_write_list, thr)


def _fwd_combine_cuda(out_partial, lse_partial, out=None, out_dtype=None):
    """lite_attention::fwd_combine = mha_combine (flash_api.cpp:1620-1720): out_partial (n, b, s, h, d) fp32,
    lse_partial (n, b, s, h) fp32 contiguous in the seqlen dimension; returns (out, softmax_lse (b, s, h) view of a
    (b, h, s) buffer).  bf16 partials are accepted as well (what LiteAttention returns)."""
    _check(out_partial.is_cuda and lse_partial.is_cuda, "out_partial / lse_partial must be CUDA tensors")
    _check(out_partial.dtype in (torch.float32, torch.bfloat16),
           "Attention combine function only support fp32 data type")                               # :1631 (+bf16)
    _check(lse_partial.dtype == torch.float32, "Attention combine function only support fp32 data type")
    _check(out_partial.dim() == 5 and lse_partial.dim() == 4, "out_partial must be 5-D, lse_partial 4-D")
    _check(out_partial.stride(-1) == 1, "Input tensor must have contiguous last dimension")
    _check(lse_partial.stride(-2) == 1, "LSE tensor must be contiguous in the seqlen dimension")
    n, b, s, h, d = out_partial.shape
    _check(n <= 8, "lite_attention::fwd_combine (sm_100a) supports at most 8 partial results")
    _check(tuple(lse_partial.shape) == (n, b, s, h), "lse_partial must be (num_splits, batch, seqlen, heads)")
    _check(d % 8 == 0, "head_dim must be a multiple of 8")
    out_type = out_dtype if out_dtype is not None else out_partial.dtype
    _check(out_type in (torch.float32, torch.bfloat16), "Output type must be FP32 or BF16")        # (fp16 not built)
    if out is not None:
        _check(out.dtype == out_type and out.device == out_partial.device and tuple(out.shape) == (b, s, h, d)
               and out.is_contiguous(), "out must be a contiguous (batch, seqlen, heads, head_dim) tensor of out_dtype")
    else:
        out = torch.empty((b, s, h, d), dtype=out_type, device=out_partial.device)
    o_parts = [out_partial[i].contiguous() for i in range(n)]
    l_parts = [lse_partial[i].transpose(1, 2).contiguous() for i in range(n)]      # (b, h, s), a view when stride(-2)==1
    lse = torch.empty((b, h, s), dtype=torch.float32, device=out_partial.device)
    if b * s > 0:
        _native.combine(o_parts, l_parts, out, lse)
    return out, lse.transpose(1, 2)


def _register():
    lib = torch.library.Library("lite_attention", "DEF")
    lib.define(_FWD_SCHEMA)
    lib.impl("fwd", _fwd_cuda, "CUDA")
    lib.define(_COMBINE_SCHEMA)
    lib.impl("fwd_combine", _fwd_combine_cuda, "CUDA")
    for name, schema in _STUB_SCHEMAS.items():
        lib.define(schema)

        def _stub(*args, _name=name, **kwargs):
            raise RuntimeError(f"lite_attention::{_name} is not supported (inference-only forward path)")
        lib.impl(name, _stub, "CompositeExplicitAutograd")
    return lib


_LIB = _register()
flash_attn_3_cuda = torch.ops.lite_attention


def maybe_contiguous(x):
    return x.contiguous() if x is not None and x.stride(-1) != 1 else x


def _flash_attn_forward(q, k, v, k_new, v_new, qv, out, cu_seqlens_q, cu_seqlens_k, cu_seqlens_k_new, seqused_q,
                        seqused_k, max_seqlen_q, max_seqlen_k, page_table, kv_batch_idx, leftpad_k, rotary_cos,
                        rotary_sin, seqlens_rotary, q_descale, k_descale, v_descale, softmax_scale, causal,
                        window_size=(-1, -1), attention_chunk=0, softcap=0.0, rotary_interleaved=True,
                        scheduler_metadata=None, num_splits=1, pack_gqa=None, sm_margin=0,
                        attn_read_list=None, attn_must_do_list=None, attn_write_list=None, thr=-3.0):
    q, k, k_new, v_new = [maybe_contiguous(x) for x in (q, k, k_new, v_new)]
    v = v.contiguous() if v.stride(-1) != 1 and v.stride(-3) != 1 else v
    out, softmax_lse, *rest = flash_attn_3_cuda.fwd(
        q, k, v, k_new, v_new, qv, out, cu_seqlens_q, cu_seqlens_k, cu_seqlens_k_new, seqused_q, seqused_k,
        max_seqlen_q, max_seqlen_k, page_table, kv_batch_idx, leftpad_k, rotary_cos, rotary_sin, seqlens_rotary,
        q_descale, k_descale, v_descale, softmax_scale, causal, window_size[0], window_size[1], attention_chunk,
        softcap, rotary_interleaved, scheduler_metadata, num_splits, pack_gqa, sm_margin,
        attn_read_list, attn_must_do_list, attn_write_list, thr=thr)
    return out, softmax_lse, *rest


class FlashAttnFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, softmax_scale, causal, qv=None, q_descale=None, k_descale=None, v_descale=None,
                window_size=(-1, -1), attention_chunk=0, softcap=0.0, num_splits=1, pack_gqa=None,
                deterministic=False, sm_margin=0, attn_read_list=None, attn_must_do_list=None,
                attn_write_list=None, thr=-3.0, return_softmax_lse=False):
        if softmax_scale is None:
            softmax_scale = (q.shape[-1] + (qv.shape[-1] if qv is not None else 0)) ** (-0.5)
        out, softmax_lse, *rest = _flash_attn_forward(
            q, k, v, None, None, qv, None, None, None, None, None, None, None, None, None, None, None, None, None,
            None, q_descale, k_descale, v_descale, softmax_scale, causal=causal, window_size=window_size,
            attention_chunk=attention_chunk, softcap=softcap, num_splits=num_splits, pack_gqa=pack_gqa,
            sm_margin=sm_margin, attn_read_list=attn_read_list, attn_must_do_list=attn_must_do_list,
            attn_write_list=attn_write_list, thr=thr)
        ctx.mark_non_differentiable(softmax_lse)
        if return_softmax_lse:
            return out, softmax_lse
        return out

    @staticmethod
    def backward(ctx, dout, *args):
        # The reference ships DISABLE_BACKWARD=TRUE (hopper/setup.py:47); run_mha_bwd throws (flash_api.cpp:1251-1254).
        raise RuntimeError("lite_attention: backward is not supported (inference-only build)")


def flash_attn_func(q, k, v, softmax_scale=None, causal=False, qv=None, q_descale=None, k_descale=None,
                    v_descale=None, window_size=(-1, -1), attention_chunk=0, softcap=0.0, num_splits=1,
                    pack_gqa=None, deterministic=False, sm_margin=0, attn_read_list=None, attn_must_do_list=None,
                    attn_write_list=None, thr=-3.0, return_softmax_lse=False):
    """Same signature and meaning as the reference's flash_attn_func (flash_attn_interface.py:547-635).

    q: (batch, seqlen_q, nheads, 128) bf16;  k, v: (batch, seqlen_k, nheads_k, 128) bf16.
    attn_read_list / attn_must_do_list / attn_write_list: int32 [>=batch, nheads, qtiles, ktiles+1] run-length skip
    lists at (128 x 176) tile granularity; thr: QK-skip threshold in the exp2 domain.
    Returns out (batch, seqlen_q, nheads, 128) [and softmax_lse (batch, nheads, seqlen_q) fp32].
    """
    return FlashAttnFunc.apply(q, k, v, softmax_scale, causal, qv, q_descale, k_descale, v_descale, window_size,
                               attention_chunk, softcap, num_splits, pack_gqa, deterministic, sm_margin,
                               attn_read_list, attn_must_do_list, attn_write_list, thr, return_softmax_lse)


def flash_attn_combine(out_partial, lse_partial, out: Optional[torch.Tensor] = None, return_lse: bool = True):
    """Merge partial attention results by their LSE (the step README.md:222-250 leaves to the caller).

    out_partial: sequence of (batch, seqlen, nheads, d) bf16 tensors, or one stacked (n, batch, seqlen, nheads, d);
    lse_partial: matching (batch, nheads, seqlen) fp32 tensors.  Returns (out, lse)."""
    o_parts = [t.contiguous() for t in out_partial]
    l_parts = [t.contiguous().float() for t in lse_partial]
    _check(len(o_parts) == len(l_parts) and 1 <= len(o_parts) <= 8, "flash_attn_combine: need 1..8 matching parts")
    b, s, h, d = o_parts[0].shape
    for o, l in zip(o_parts, l_parts):
        _check(o.is_cuda and o.dtype == torch.bfloat16 and o.shape == (b, s, h, d), "flash_attn_combine: bad out part")
        _check(l.shape == (b, h, s), "flash_attn_combine: bad lse part")
    dev = o_parts[0].device
    _check(all(o.device == dev for o in o_parts) and all(l.device == dev for l in l_parts),
           "flash_attn_combine: all parts must be on one device")
    if out is None:
        out = torch.empty_like(o_parts[0])
    else:
        _check(out.dtype in (torch.bfloat16, torch.float32) and tuple(out.shape) == (b, s, h, d)
               and out.is_contiguous() and out.device == dev,
               "flash_attn_combine: out must be a contiguous (batch, seqlen, heads, d) bf16/fp32 tensor on the parts' device")
    lse = torch.empty((b, h, s), dtype=torch.float32, device=out.device) if return_lse else None
    _native.combine(o_parts, l_parts, out, lse)
    return (out, lse) if return_lse else out
