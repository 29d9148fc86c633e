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
