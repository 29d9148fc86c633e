"""ctypes binding of ``include/pfpn_b200.h`` -- the only way Python reaches the kernels.

There is deliberately NO fallback: if ``libpfpn_b200.so`` is absent the import
raises, and every wrapper raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libpfpn_b200.so"

HEAD_FWD, HEAD_GRAD, HEAD_PPO = 0, 1, 2
HEAD_FLAG_TANH = 1


class PfpnError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[pfpn status {status}] {msg}")
        self.status = status


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} not found: build it with `python pfpn_b200/build.py` "
            "(or __graft_entry__.build()). pfpn_b200 has no CPU / PyTorch fallback.")
    return C.CDLL(str(_LIB_PATH))


lib = _load()

_f32p = C.c_void_p  # raw device pointers travel as integers


class HeadArgs(C.Structure):
    """Mirror of ``pfpn_head_args`` (field order and types must match the header)."""
    _fields_ = [
        ("logits", _f32p), ("loc", _f32p), ("logstd", _f32p), ("value", _f32p),
        ("g_lp", _f32p), ("g_ent_ba", _f32p), ("g_ent", C.c_float),
        ("adv", _f32p), ("lp_old", _f32p), ("adv_stats", _f32p),
        ("eps_clip", C.c_float), ("loss_scale", C.c_float),
        ("lp", _f32p), ("ent", _f32p), ("ent_ba", _f32p), ("dlogits", _f32p),
        ("dloc", _f32p), ("dlogstd", _f32p), ("dvalue", _f32p), ("loss", _f32p),
        ("B", C.c_int32), ("A", C.c_int32), ("P", C.c_int32),
        ("mode", C.c_uint32), ("flags", C.c_uint32),
    ]


class HeadPush(C.Structure):
    """Mirror of ``pfpn_head_push``."""
    _fields_ = [("out", C.c_void_p * 8), ("flags", C.c_void_p * 8), ("ticket", C.c_void_p),
                ("nranks", C.c_int32), ("value", C.c_int32),
                ("protocol", C.c_int32), ("consume_value", C.c_int32), ("consume_rows", C.c_void_p),
                ("consume_out", C.c_void_p), ("consume_scale", C.c_float), ("reserved", C.c_int32)]


class RolloutArgs(C.Structure):
    """Mirror of ``pfpn_rollout_args``."""
    _fields_ = [
        ("logits", _f32p), ("loc", _f32p), ("logstd", _f32p), ("ext_uniform", C.c_void_p), ("ext_normal", _f32p),
        ("action", _f32p), ("idx", C.c_void_p), ("lp", _f32p), ("ent", _f32p),
        ("max_active", _f32p), ("sum_active", _f32p),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("B", C.c_int32), ("A", C.c_int32), ("P", C.c_int32),
    ]


class SacHeadArgs(C.Structure):
    """Mirror of ``pfpn_sac_head_args``."""
    _fields_ = [
        ("logits", _f32p), ("loc", _f32p), ("logstd", _f32p), ("ext_uniform", _f32p), ("ext_normal", _f32p),
        ("g_sample", _f32p), ("g_lp", _f32p),
        ("sample", _f32p), ("s_pre", _f32p), ("idx", C.c_void_p), ("logp", _f32p),
        ("dlogits", _f32p), ("dloc", _f32p), ("dlogstd", _f32p),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("B", C.c_int32), ("A", C.c_int32), ("P", C.c_int32), ("reserved", C.c_int32), ("offset_dev", C.c_void_p),
    ]


class SyncArgs(C.Structure):
    """Mirror of ``pfpn_sync_args``."""
    _fields_ = [
        ("grads", C.c_void_p), ("n_params", C.c_size_t), ("n_total", C.c_size_t), ("clip", C.c_float),
        ("new_mean", C.c_void_p), ("new_std", C.c_void_p), ("state_mean", C.c_void_p), ("state_std", C.c_void_p),
        ("max_active", C.c_void_p), ("sum_active", C.c_void_p), ("S", C.c_int32), ("AP", C.c_int32),
        ("params", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("counters", C.c_void_p), ("norm_scale", C.c_void_p), ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t),
        ("stage", C.c_void_p), ("reduced", C.c_void_p), ("flags", C.c_void_p),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("two_phase", C.c_int32),
        ("params_lo", C.c_void_p),
    ]


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


pfpn_abi_version = _sig("pfpn_abi_version", C.c_int, [])
pfpn_status_string = _sig("pfpn_status_string", C.c_char_p, [C.c_int])
pfpn_head_workspace_bytes = _sig("pfpn_head_workspace_bytes", C.c_int,
                                 [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)])
pfpn_head_logprob = _sig("pfpn_head_logprob", C.c_int,
                         [C.POINTER(HeadArgs), C.c_void_p, C.c_size_t, C.c_void_p])
pfpn_adv_stats = _sig("pfpn_adv_stats", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p])
pfpn_head_launch_info = _sig("pfpn_head_launch_info", C.c_int,
                             [C.c_int32, C.c_int32, C.c_uint32, C.POINTER(C.c_int32)])


def check(status: int) -> None:
    if status != 0:
        raise PfpnError(status, pfpn_status_string(status).decode())


def exported_symbols():
    """Names declared in include/pfpn_b200.h (used by the CPU test-suite)."""
    import re
    hdr = (Path(__file__).resolve().parent.parent / "include" / "pfpn_b200.h").read_text()
    return sorted(set(re.findall(r"\b(pfpn_[a-z0-9_]+)\s*\(", hdr)))


# ---- K2 / K3 / K4 / K5 ------------------------------------------------------------------------
RESAMPLE_FLAG_TANH = 1


class SampleArgs(C.Structure):
    """Mirror of ``pfpn_sample_args``."""
    _fields_ = [
        ("logits", _f32p), ("loc", _f32p), ("logstd", _f32p), ("ext_uniform", C.c_void_p),
        ("ext_normal", _f32p), ("action", _f32p), ("idx", C.c_void_p),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("B", C.c_int32), ("A", C.c_int32), ("P", C.c_int32),
    ]


class RSampleArgs(C.Structure):
    """Mirror of ``pfpn_rsample_args``."""
    _fields_ = [
        ("logits", _f32p), ("loc", _f32p), ("logstd", _f32p), ("ext_uniform", _f32p), ("ext_normal", _f32p),
        ("sample", _f32p), ("s_pre", _f32p), ("idx", C.c_void_p),
        ("g_sample", _f32p), ("g_s_pre", _f32p), ("dlogits", _f32p), ("dloc", _f32p), ("dlogstd", _f32p),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("B", C.c_int32), ("A", C.c_int32), ("P", C.c_int32), ("reserved", C.c_int32), ("offset_dev", C.c_void_p),
    ]


class ResampleArgs(C.Structure):
    """Mirror of ``pfpn_resample_args``."""
    _fields_ = [
        ("max_active", _f32p), ("sum_active", _f32p), ("loc", _f32p), ("logstd", _f32p),
        ("bias", _f32p), ("weight", _f32p),
        ("ext_cat_u", C.c_void_p), ("ext_choice", C.c_void_p), ("ext_noise_u", _f32p),
        ("out_M", C.c_void_p), ("out_nuniq", C.c_void_p), ("out_invalid", C.c_void_p), ("out_cand", C.c_void_p),
        ("out_src", C.c_void_p), ("out_col", C.c_void_p), ("out_tcol", C.c_void_p), ("out_uniq", C.c_void_p),
        ("out_idx", C.c_void_p), ("out_count", C.c_void_p), ("out_delta", C.c_void_p),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("threshold", C.c_float),
        ("A", C.c_int32), ("P", C.c_int32), ("H", C.c_int32),
        ("resample", C.c_int32), ("flags", C.c_uint32),
    ]


pfpn_head_sample = _sig("pfpn_head_sample", C.c_int, [C.POINTER(SampleArgs), C.c_void_p])
pfpn_head_rsample_fwd = _sig("pfpn_head_rsample_fwd", C.c_int, [C.POINTER(RSampleArgs), C.c_void_p])
pfpn_rsample_bwd_workspace_bytes = _sig("pfpn_rsample_bwd_workspace_bytes", C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)])
pfpn_head_rsample_bwd = _sig("pfpn_head_rsample_bwd", C.c_int, [C.POINTER(RSampleArgs), C.c_void_p, C.c_size_t, C.c_void_p])
pfpn_rollout_workspace_bytes = _sig("pfpn_rollout_workspace_bytes", C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)])
pfpn_head_rollout = _sig("pfpn_head_rollout", C.c_int, [C.POINTER(RolloutArgs), C.c_void_p, C.c_size_t, C.c_void_p])
pfpn_sac_head_workspace_bytes = _sig("pfpn_sac_head_workspace_bytes", C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)])
pfpn_sac_head_fwd_bwd = _sig("pfpn_sac_head_fwd_bwd", C.c_int, [C.POINTER(SacHeadArgs), C.c_void_p, C.c_size_t, C.c_void_p])
pfpn_head_finalize_partials = _sig("pfpn_head_finalize_partials", C.c_int,
                                   [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p])
pfpn_head_mean = _sig("pfpn_head_mean", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                                  C.c_int32, C.c_int32, C.c_uint32, C.c_void_p])
pfpn_stats_workspace_bytes = _sig("pfpn_stats_workspace_bytes", C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)])
pfpn_stats_update = _sig("pfpn_stats_update", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                        C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p])
pfpn_resample_workspace_bytes = _sig("pfpn_resample_workspace_bytes", C.c_int,
                                     [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)])
pfpn_resample = _sig("pfpn_resample", C.c_int, [C.POINTER(ResampleArgs), C.c_void_p, C.c_size_t, C.c_void_p])


# ---- K6 / K7 ---------------------------------------------------------------------------------------
_vp, _i32, _f = C.c_void_p, C.c_int32, C.c_float
pfpn_mlp_linear_fwd = _sig("pfpn_mlp_linear_fwd", C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp])
pfpn_mlp_linear_bwd_input = _sig("pfpn_mlp_linear_bwd_input", C.c_int,
                                 [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp])
pfpn_mlp_wgrad_workspace_bytes = _sig("pfpn_mlp_wgrad_workspace_bytes", C.c_int,
                                      [_i32, _i32, _i32, C.POINTER(C.c_size_t)])
pfpn_mlp_linear_bwd_weight = _sig("pfpn_mlp_linear_bwd_weight", C.c_int,
                                  [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, C.c_size_t, _vp])
pfpn_state_normalize = _sig("pfpn_state_normalize", C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f, _i32, _vp])
pfpn_normalizer_update = _sig("pfpn_normalizer_update", C.c_int, [_vp, _vp, _vp, _i32, _i32, _f, _vp, _vp])
pfpn_normalizer_scratch_bytes = _sig("pfpn_normalizer_scratch_bytes", C.c_int, [_i32, C.POINTER(C.c_size_t)])
pfpn_normalizer_update_dev = _sig("pfpn_normalizer_update_dev", C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, C.c_size_t, _vp])
pfpn_value_loss = _sig("pfpn_value_loss", C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _f, _f, _vp])
pfpn_gae = _sig("pfpn_gae", C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _f, _f, _vp])
pfpn_sac_losses = _sig("pfpn_sac_losses", C.c_int, [_vp] * 11 + [_f, _f, _f, _i32] + [_vp] * 6 + [_vp])
pfpn_axpby = _sig("pfpn_axpby", C.c_int, [_vp, _vp, C.c_size_t, _f, _f, _vp])
pfpn_clip_by_global_norm = _sig("pfpn_clip_by_global_norm", C.c_int, [_vp, C.c_size_t, _f, _vp, _vp, C.c_size_t, _vp])
pfpn_adam_step = _sig("pfpn_adam_step", C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _f, _f, _f, _f, C.c_int64, _f, _vp])
pfpn_adam_step_dev = _sig("pfpn_adam_step_dev", C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _f, _f, _f, _f, _vp, _i32, _f, _vp])
pfpn_tc_gemm_nt = _sig("pfpn_tc_gemm_nt", C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp])
pfpn_tc_gemm_nn = _sig("pfpn_tc_gemm_nn", C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp])
pfpn_tc_gemm_nt_lo = _sig("pfpn_tc_gemm_nt_lo", C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp])
pfpn_tc_gemm_nn_lo = _sig("pfpn_tc_gemm_nn_lo", C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp])
pfpn_split_lo = _sig("pfpn_split_lo", C.c_int, [_vp, _vp, C.c_size_t, _vp])
pfpn_transpose = _sig("pfpn_transpose", C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _vp])
pfpn_tc_wgrad_workspace_bytes = _sig("pfpn_tc_wgrad_workspace_bytes", C.c_int, [_i32, _i32, _i32, C.POINTER(C.c_size_t)])
pfpn_tc_linear_bwd_weight = _sig("pfpn_tc_linear_bwd_weight", C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, C.c_size_t, _vp])
pfpn_bias_grad = _sig("pfpn_bias_grad", C.c_int, [_vp, _i32, _vp, _i32, _i32, _vp, C.c_size_t, _vp])
pfpn_enable_peer_access = _sig("pfpn_enable_peer_access", C.c_int, [_i32])
pfpn_peer_signal = _sig("pfpn_peer_signal", C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp])
pfpn_peer_allreduce_adam = _sig("pfpn_peer_allreduce_adam", C.c_int,
                                [_vp, _vp, _i32, _i32, _i32, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, _f, _f, _f, _f,
                                 C.c_int64, _vp])
pfpn_peer_allreduce_adam_rs = _sig("pfpn_peer_allreduce_adam_rs", C.c_int,
                                   [_vp, _vp, _vp, _i32, _i32, _i32, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, _f, _f, _f, _f,
                                    C.c_int64, _vp])
pfpn_peer_allreduce_sum = _sig("pfpn_peer_allreduce_sum", C.c_int, [_vp, _vp, _i32, _i32, _i32, C.c_size_t, _vp, _f, _vp])
pfpn_head_logprob_push = _sig("pfpn_head_logprob_push", C.c_int,
                              [C.POINTER(HeadArgs), C.c_void_p, C.c_size_t, C.POINTER(HeadPush), C.c_void_p])
pfpn_peer_gather_sum = _sig("pfpn_peer_gather_sum", C.c_int, [_vp, _vp, _i32, _i32, C.c_size_t, _vp, _f, _vp])
pfpn_peer_gather_sum_packets = _sig("pfpn_peer_gather_sum_packets", C.c_int, [_vp, _i32, _i32, C.c_size_t, _vp, _f, _vp])
pfpn_sync_step_scratch_bytes = _sig("pfpn_sync_step_scratch_bytes", C.c_int, [C.POINTER(C.c_size_t)])
pfpn_sync_step = _sig("pfpn_sync_step", C.c_int, [C.POINTER(SyncArgs), C.c_void_p])
pfpn_peer_alloc = _sig("pfpn_peer_alloc", C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p])
pfpn_peer_open = _sig("pfpn_peer_open", C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)])
pfpn_peer_close = _sig("pfpn_peer_close", C.c_int, [_vp])
pfpn_peer_free = _sig("pfpn_peer_free", C.c_int, [_vp])
