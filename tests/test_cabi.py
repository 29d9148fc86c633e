"""CPU checks of the drop-in boundary: the library loads and exports every symbol
that include/pfpn_b200.h declares; argument validation works without a GPU."""
import ctypes as C

from pfpn_b200 import _cabi


def test_library_exports_every_declared_symbol():
    names = _cabi.exported_symbols()
    assert "pfpn_head_logprob" in names and len(names) >= 6
    for n in names:
        assert hasattr(_cabi.lib, n), n


def test_abi_version_and_status_strings():
    assert _cabi.pfpn_abi_version() == 2
    assert _cabi.pfpn_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert _cabi.pfpn_status_string(code).startswith(b"pfpn:")


def test_argument_errors_without_gpu():
    n = C.c_size_t(0)
    assert _cabi.pfpn_head_workspace_bytes(36, 35, C.byref(n)) == 0 and n.value > 0
    assert _cabi.pfpn_head_workspace_bytes(0, 35, C.byref(n)) == -1
    assert _cabi.pfpn_head_logprob(None, None, 0, None) == -1
    a = _cabi.HeadArgs()
    a.B, a.A, a.P = 4, 36, 35
    assert _cabi.pfpn_head_logprob(C.byref(a), None, 0, None) == -1  # null pointers


def test_head_args_layout_matches_header():
    # 20 pointers/floats interleaved + 5 ints: catch accidental reordering
    assert _cabi.HeadArgs.logits.offset == 0
    assert _cabi.HeadArgs.g_ent.offset == 48 and _cabi.HeadArgs.adv.offset == 56
    assert _cabi.HeadArgs.eps_clip.offset == 80 and _cabi.HeadArgs.loss_scale.offset == 84
    assert _cabi.HeadArgs.lp.offset == 88 and _cabi.HeadArgs.loss.offset == 144
    assert _cabi.HeadArgs.B.offset == 152 and C.sizeof(_cabi.HeadArgs) == 176


def test_every_header_function_has_a_ctypes_signature():
    """The host mirror binds the whole boundary: each function of include/pfpn_b200.h has a typed ctypes entry."""
    missing = []
    for n in _cabi.exported_symbols():
        f = getattr(_cabi, n, None)
        if f is None or getattr(f, "argtypes", None) is None:
            missing.append(n)
    assert missing == []


def test_other_struct_layouts_match_the_static_asserts_of_capi_cu():
    assert C.sizeof(_cabi.SampleArgs) == 88 and C.sizeof(_cabi.RSampleArgs) == 144 and C.sizeof(_cabi.ResampleArgs) == 200
    assert _cabi.RSampleArgs.offset_dev.offset == 136
    assert C.sizeof(_cabi.SacHeadArgs) == 152 and _cabi.SacHeadArgs.offset_dev.offset == 144
    assert C.sizeof(_cabi.HeadPush) == 176 and _cabi.HeadPush.consume_rows.offset == 152 and _cabi.HeadPush.consume_scale.offset == 168


def test_argument_errors_of_the_widened_entry_points_without_gpu():
    n = C.c_size_t(0)
    assert _cabi.pfpn_gae(None, None, None, None, 4, 8, 0.99, 0.94, None) == -1
    assert _cabi.pfpn_axpby(None, None, 16, 0.5, 0.5, None) == -1
    assert _cabi.pfpn_sac_losses(*([None] * 11), 0.95, 0.5, -36.0, 8, *([None] * 6), None) == -1
    assert _cabi.pfpn_stats_workspace_bytes(4096, 36, 35, C.byref(n)) == 0 and n.value > 0
    assert _cabi.pfpn_stats_update(None, None, None, None, 8, 36, 35, None, 0, None) == -1
    assert _cabi.pfpn_rsample_bwd_workspace_bytes(4096, 36, 100, C.byref(n)) == 0 and n.value >= 592 * 2 * 3600 * 4
    assert _cabi.pfpn_tc_wgrad_workspace_bytes(65536, 1024, 512, C.byref(n)) == 0 and n.value >= 32 * (1024 * 512 + 512) * 4
    assert _cabi.pfpn_tc_wgrad_workspace_bytes(8192, 512, 128, C.byref(n)) == 0  # small batch: finer split-K, larger scratch per row
    assert n.value >= 8 * (512 * 128 + 128) * 4
    assert _cabi.pfpn_peer_allreduce_sum(None, None, 0, 2, 1, 16, None, 1.0, None) == -1
