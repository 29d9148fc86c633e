"""Wrapper over K5 (``pfpn_resample``): in-place dead-particle resampling
(/root/reference/networks/actor_critic/a2c.py:385-474)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _cabi
from .head import _stream_ptr

_ws: dict = {}


def _chk(t: torch.Tensor, name: str, shape):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
        raise ValueError(f"{name}: expected contiguous float32 CUDA tensor of shape {tuple(shape)}")


def resample_(max_active, sum_active, loc, logstd, bias, weight, *, resample: int = -1,
              threshold: Optional[float] = None, tanh: bool = False, seed: int = 0, offset: int = 0,
              ext_cat_u=None, ext_choice=None, ext_noise_u=None, verify: bool = False) -> Optional[Dict]:
    """Updates loc / logstd / bias / weight in place and zeroes the statistics.  With
    ``verify=True`` returns the integer decisions (device tensors) for parity checks."""
    A, P = max_active.shape
    H = 0 if weight is None else weight.shape[0]
    for t, n in ((max_active, "max_active"), (sum_active, "sum_active"), (loc, "loc"), (logstd, "logstd")):
        _chk(t, n, (A, P))
    _chk(bias, "bias", (A * P,))
    if weight is not None:
        _chk(weight, "weight", (H, A * P))
    dev = max_active.device
    a = _cabi.ResampleArgs()
    a.max_active, a.sum_active, a.loc, a.logstd = (max_active.data_ptr(), sum_active.data_ptr(), loc.data_ptr(),
                                                  logstd.data_ptr())
    a.bias = bias.data_ptr()
    a.weight = None if weight is None else weight.data_ptr()
    keep = []
    if ext_cat_u is not None:
        assert ext_cat_u.dtype == torch.float64 and tuple(ext_cat_u.shape) == (A, P)
        ext_cat_u = ext_cat_u.contiguous()
        keep.append(ext_cat_u)
        a.ext_cat_u = ext_cat_u.data_ptr()
    if ext_choice is not None:
        assert ext_choice.dtype == torch.int32 and ext_choice.numel() >= A * P
        keep.append(ext_choice)
        a.ext_choice = ext_choice.data_ptr()
    if ext_noise_u is not None:
        assert ext_noise_u.dtype == torch.float32 and ext_noise_u.numel() >= A * P
        keep.append(ext_noise_u)
        a.ext_noise_u = ext_noise_u.data_ptr()
    out = None
    if verify:
        i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        out = dict(M=i32(1), nuniq=i32(1), invalid=i32(A * P, 2), cand=i32(A, P if resample < 0 else min(P, resample)),
                   src=i32(A * P), col=i32(A * P), tcol=i32(A * P), uniq=i32(A * P), idx=i32(A * P),
                   count=i32(A * P), delta=i32(A * P))
        a.out_M, a.out_nuniq, a.out_invalid, a.out_cand = (out["M"].data_ptr(), out["nuniq"].data_ptr(),
                                                           out["invalid"].data_ptr(), out["cand"].data_ptr())
        a.out_src, a.out_col, a.out_tcol = out["src"].data_ptr(), out["col"].data_ptr(), out["tcol"].data_ptr()
        a.out_uniq, a.out_idx, a.out_count, a.out_delta = (out["uniq"].data_ptr(), out["idx"].data_ptr(),
                                                           out["count"].data_ptr(), out["delta"].data_ptr())
    a.seed, a.offset = seed, offset
    a.threshold = float(threshold) if threshold else 0.0
    a.A, a.P, a.H, a.resample = A, P, H, resample
    a.flags = _cabi.RESAMPLE_FLAG_TANH if tanh else 0
    n = C.c_size_t(0)
    _cabi.check(_cabi.pfpn_resample_workspace_bytes(A, P, H, C.byref(n)))
    key = (dev.index, n.value)
    ws = _ws.get(key)
    if ws is None:
        ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
        _ws[key] = ws
    with torch.cuda.device(dev):
        _cabi.check(_cabi.pfpn_resample(C.byref(a), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return out
