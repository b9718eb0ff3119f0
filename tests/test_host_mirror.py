"""CPU: the host-side mirror of hopper/lite_attention.py against fixtures generated from the reference itself
(oracle/gen_golden.py), the error behaviour, and the C-ABI library's exported symbols (no compute calls)."""
import ctypes
import json
import os
import re

import pytest
import torch

import liteattention_b200
from liteattention_b200 import LiteAttention, SeqParallelLiteAttention
from liteattention_b200 import lite_attention as la_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def book(golden_dir):
    return json.load(open(os.path.join(golden_dir, "host_bookkeeping.json")))


def test_dropin_import_surface():
    import lite_attention
    from lite_attention import LiteAttention as LA2, SeqParallelLiteAttention as SP2
    from lite_attention._internal.flash_attn_interface import flash_attn_func
    import lite_attention._C  # noqa: F401
    assert LA2 is LiteAttention and SP2 is SeqParallelLiteAttention
    assert lite_attention.__version__ == "0.2.0" == liteattention_b200.__version__
    assert callable(flash_attn_func)
    schema = str(torch.ops.lite_attention.fwd.default._schema)
    for frag in ("Tensor? attn_read_list=None", "Tensor? attn_must_do_list=None", "Tensor? attn_write_list=None",
                 "float thr=-3.", "Tensor(out!)? out=None"):
        assert frag in schema, schema
    assert len(torch.ops.lite_attention.fwd.default._schema.arguments) == 38


def test_get_mn(book):
    for e in book["get_MN"]:
        assert list(LiteAttention.get_MN(e["head_dim"], e["element_size"], e["v_colmajor"])) == e["mn"]


def test_init_skip_list(book):
    for e in book["init_skip_list"]:
        b, s, h, d = e["args"]
        sl = LiteAttention.init_skip_list(b, s, h, d, False, torch.bfloat16, "cpu")
        assert list(sl.shape) == e["shape"] and sl.dtype == torch.int32
        assert sl[0, 0, 0, 0, :4].tolist() == e["row_prefix"]
        assert bool((sl == sl[0, 0, 0, 0]).all()) == e["all_rows_equal"]
        assert int((sl != 0).sum()) == e["nonzero"]


def test_expand_must_do(book):
    for e in book["expand_must_do"]:
        s = e["seq_len"]
        q = torch.zeros(1, s, 2, 128, dtype=torch.bfloat16)
        shape = tuple(e["shape"])
        ex = LiteAttention._expand_must_do_list(list(e["must_do_list"]), shape, q, q)
        assert list(ex.shape) == e["shape"] and ex.dtype == torch.int32 and ex.is_contiguous()
        assert ex[0, 0, 0, :len(e["row_prefix"])].tolist() == e["row_prefix"]
        assert bool((ex == ex[0, 0, 0]).all())


def test_calc_percentage_keeps_reference_formula(book):
    for e in book["calc_percentage"]:
        if "args" in e:
            s, h = e["args"]
            sl = LiteAttention.init_skip_list(1, s, h, 128, False, torch.bfloat16, "cpu")
            assert float(LiteAttention.calc_percentage(sl[0])) == pytest.approx(e["value"], rel=1e-6)
            assert LiteAttention.sparsity(sl[0]) == pytest.approx(0.0, abs=1e-12)   # the correct figure
        else:
            row = torch.tensor(e["row"], dtype=torch.int32).view(1, 1, 1, -1)
            assert float(LiteAttention.calc_percentage(row)) == pytest.approx(e["value"], rel=1e-6)
            # [6, 11,8, 5,3, 2,0] lists 4+3+3 = 10 of 12 tiles
            assert LiteAttention.sparsity(row) == pytest.approx(1 - 10 / 12)


def test_state_machine_matches_reference_trace(book, monkeypatch):
    calls = []

    def fake(**kw):
        calls.append(kw)
        return kw["q"]
    monkeypatch.setattr(la_mod, "flash_attn_func", fake)
    la = LiteAttention(enable_skipping=True, threshold=-10.0, max_batch_size=2)
    q = torch.zeros(1, 1000, 2, 128, dtype=torch.bfloat16)
    for ref in book["state_trace"][:3]:
        la(q, q, q)
        kw = calls[-1]
        assert la._phase == ref["phase_after"]
        assert int(kw["attn_read_list"].data_ptr() == la._skip_list[1].data_ptr()) == ref["read_is_buf"]
        assert int(kw["attn_write_list"].data_ptr() == la._skip_list[1].data_ptr()) == ref["write_is_buf"]
        assert kw["thr"] == ref["thr"]
        # documented divergence: the default must-do list ([2,0,0] rows in the reference) is passed as None
        assert kw["attn_must_do_list"] is None and ref["must_do_prefix"] == [2, 0, 0]
    q2 = torch.zeros(1, 1500, 2, 128, dtype=torch.bfloat16)
    la(q2, q2, q2)
    ref = book["state_trace"][3]
    assert la._phase == ref["after_shape_change_phase"] and list(la._skip_list.shape) == ref["skip_list_shape"]
    # explicit must-do list: expanded like the reference, cached across calls
    la(q2, q2, q2, must_do_list=[1499, 0])
    md = calls[-1]["attn_must_do_list"]
    assert md.shape == la._skip_list.shape[1:] and md[0, 0, 0, :3].tolist() == [2, 9, 0]
    la(q2, q2, q2, must_do_list=[1499, 0])
    assert calls[-1]["attn_must_do_list"] is md
    # reset
    la.reset_skip_state()
    assert la._skip_list is None and la._phase == 0


def test_disabled_skipping_passes_no_lists(monkeypatch):
    calls = []
    monkeypatch.setattr(la_mod, "flash_attn_func", lambda **kw: calls.append(kw) or kw["q"])
    la = LiteAttention(enable_skipping=False)
    q = torch.zeros(1, 300, 2, 128, dtype=torch.bfloat16)
    la(q, q, q)          # the reference raises AttributeError here (lite_attention.py:262 [BUG]); intent = dense
    kw = calls[-1]
    assert kw["attn_read_list"] is None and kw["attn_write_list"] is None and kw["attn_must_do_list"] is None
    la.enable_skip_optimization(True)
    la(q, q, q)
    assert calls[-1]["attn_read_list"] is not None


def test_error_behaviour(book):
    la = LiteAttention()
    with pytest.raises(ValueError) as ei:
        la.set_threshold(0.5)
    assert str(ei.value) == book["errors"]["set_threshold_positive"]
    with pytest.raises(ValueError):
        LiteAttention(threshold=0.0)
    os.environ["LITE_ATTENTION_DEBUG"] = "TRUE"
    try:
        LiteAttention(threshold=2.0)
    finally:
        del os.environ["LITE_ATTENTION_DEBUG"]
    q3 = torch.zeros(3, 1000, 2, 128, dtype=torch.bfloat16)
    with pytest.raises(AssertionError) as ei:
        LiteAttention(max_batch_size=2)._get_read_write_lists(q3, q3)
    assert str(ei.value) == book["errors"]["batch_gt_max"]


def test_seq_parallel(book):
    sp = SeqParallelLiteAttention(3, True, -5.0, 2)
    assert sp.num_nodes == book["seq_parallel"]["num_nodes"]
    assert [x.threshold for x in sp.lite_attention] == book["seq_parallel"]["thresholds"]
    sp.set_threshold(-2.0)
    assert all(x.threshold == -2.0 for x in sp.lite_attention)
    sp.enable_skip_optimization(False)
    assert not any(x.enable_skipping for x in sp.lite_attention)
    q = torch.zeros(1, 10, 1, 128, dtype=torch.bfloat16)
    with pytest.raises(AssertionError):
        sp(q, q, q, split_idx=3)
    sp = SeqParallelLiteAttention(2, compact_state=True)                # keyword-only extension reaches every split
    assert all(x._compact for x in sp.lite_attention)


def test_functional_api_has_no_cpu_fallback():
    from liteattention_b200 import flash_attn_func
    q = torch.zeros(1, 256, 2, 128, dtype=torch.bfloat16)
    with pytest.raises(NotImplementedError):
        flash_attn_func(q, q, q)            # CPU tensors: the dispatcher has no CPU kernel and nothing falls back


def test_c_abi_exports_every_declared_symbol(native_lib):
    hdr = open(os.path.join(ROOT, "include", "liteattn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(la_[a-z0-9_]+)\s*\(", hdr))
    assert {"la_fwd_sm100", "la_skip_update_sm100", "la_fwd_skip_sm100", "la_combine_sm100", "la_get_tile_mn",
            "la_rope_cast_sm100", "la_last_error", "la_abi_version", "la_launch_count"} <= declared
    lib = ctypes.CDLL(native_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/liteattn_b200.h but not exported"
    assert native_lib.lib().la_abi_version() == int(re.search(r"#define LA_ABI_VERSION (\d+)", hdr).group(1))
    assert native_lib.get_tile_mn(128) == (128, 176, True)
    assert native_lib.get_tile_mn(64) == (192, 192, False)      # table matches get_MN; kernel not built -> unsupported
    # struct layouts agree with the header (sizes computed by hand from the C declaration)
    assert ctypes.sizeof(native_lib.FwdParams) == 5 * 8 + 12 * 8 + 6 * 4 + 4 + 4 + 2 * 8 + 4 * 4 + 8 * 8   # + out_is_f32, n_out_peers, out_rows_per_peer, reserved, out_peer[8]
    assert ctypes.sizeof(native_lib.RopeParams) == 4 * 8 + 3 * 8 + 6 * 4
    assert ctypes.sizeof(native_lib.UpdateParams) == 4 * 8 + 4 * 4 + 4 + 4 + 8


def test_host_head_group_schedule():
    """Head groups of the host-resident call (liteattention_b200/lite_attention.py:host_head_groups): a partition of the
    heads with a small first upload and a small last download; one big group + the small last one while another call is
    in flight; LITE_ATTENTION_HOST_CHUNKS overrides and is validated."""
    from liteattention_b200.lite_attention import host_head_groups
    for heads in (1, 2, 3, 4, 5, 8, 12, 16, 24, 40, 64, 96):
        g = host_head_groups(heads)
        assert sum(g) == heads and min(g) >= 1
        gb = host_head_groups(heads, busy=True)
        assert sum(gb) == heads and min(gb) >= 1 and len(gb) <= 2
        if heads >= 20:
            assert g[0] <= heads // 10 and g[-1] <= heads // 8
    assert host_head_groups(40) == [2, 4, 8, 13, 9, 4] and host_head_groups(40, True) == [36, 4]
    import os
    os.environ["LITE_ATTENTION_HOST_CHUNKS"] = "10,10,20"
    try:
        assert host_head_groups(40) == [10, 10, 20]
        with pytest.raises(ValueError):
            host_head_groups(32)
    finally:
        del os.environ["LITE_ATTENTION_HOST_CHUNKS"]
