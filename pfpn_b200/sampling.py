"""Wrappers over K2 (plain sample), K3 (reparameterised sample fwd/bwd), the deterministic
action and K4 (activity statistics).  torch tensors carry device memory only."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _cabi
from .head import _f32c, _stream_ptr


def _i32(*shape, device):
    return torch.empty(*shape, dtype=torch.int32, device=device)


def sample_plain(logits, loc, logstd, *, seed: int = 0, offset: int = 0, ext_uniform: Optional[torch.Tensor] = None,
                 ext_normal: Optional[torch.Tensor] = None):
    """utils.py:187-194.  Returns (action [B,A] f32, idx [B,A] int32)."""
    logits, loc, logstd = _f32c(logits, "logits"), _f32c(loc, "loc"), _f32c(logstd, "logstd")
    B, A, P = logits.shape
    a = _cabi.SampleArgs()
    a.logits, a.loc, a.logstd = logits.data_ptr(), loc.data_ptr(), logstd.data_ptr()
    keep = []
    if ext_uniform is not None:
        if ext_uniform.dtype != torch.float64 or ext_uniform.shape != (B, A):
            raise ValueError("ext_uniform must be float64 [B, A]")
        ext_uniform = ext_uniform.contiguous()
        keep.append(ext_uniform)
        a.ext_uniform = ext_uniform.data_ptr()
    if ext_normal is not None:
        ext_normal = _f32c(ext_normal, "ext_normal")
        if ext_normal.shape != (B, A, P):
            raise ValueError("ext_normal must be float32 [B, A, P]")
        keep.append(ext_normal)
        a.ext_normal = ext_normal.data_ptr()
    action = torch.empty(B, A, dtype=torch.float32, device=logits.device)
    idx = _i32(B, A, device=logits.device)
    a.action, a.idx = action.data_ptr(), idx.data_ptr()
    a.seed, a.offset, a.B, a.A, a.P = seed, offset, B, A, P
    with torch.cuda.device(logits.device):
        _cabi.check(_cabi.pfpn_head_sample(C.byref(a), _stream_ptr()))
    return action, idx


def _rsample_args(logits, loc, logstd, seed, offset, ext_uniform, ext_normal, offset_dev=None):
    logits, loc, logstd = _f32c(logits, "logits"), _f32c(loc, "loc"), _f32c(logstd, "logstd")
    B, A, P = logits.shape
    a = _cabi.RSampleArgs()
    a.logits, a.loc, a.logstd = logits.data_ptr(), loc.data_ptr(), logstd.data_ptr()
    keep = [logits, loc, logstd]
    if (ext_uniform is None) != (ext_normal is None):
        raise ValueError("pass both ext_uniform and ext_normal or neither")
    if ext_uniform is not None:
        ext_uniform, ext_normal = _f32c(ext_uniform, "ext_uniform"), _f32c(ext_normal, "ext_normal")
        keep += [ext_uniform, ext_normal]
        a.ext_uniform, a.ext_normal = ext_uniform.data_ptr(), ext_normal.data_ptr()
    a.seed, a.offset, a.B, a.A, a.P = seed, offset, B, A, P
    if offset_dev is not None:  # int64 device word added to `offset` by the kernel (graph replays draw fresh variates)
        assert offset_dev.dtype == torch.int64 and offset_dev.is_cuda
        a.offset_dev = offset_dev.data_ptr()
        keep.append(offset_dev)
    return a, keep


def rsample_fwd(logits, loc, logstd, *, seed=0, offset=0, ext_uniform=None, ext_normal=None, offset_dev=None):
    """utils.py:156-186 forward.  Returns (sample = tanh(s_), s_, idx)."""
    a, keep = _rsample_args(logits, loc, logstd, seed, offset, ext_uniform, ext_normal, offset_dev)
    dev = logits.device
    sample = torch.empty(a.B, a.A, dtype=torch.float32, device=dev)
    s_pre = torch.empty_like(sample)
    idx = _i32(a.B, a.A, device=dev)
    a.sample, a.s_pre, a.idx = sample.data_ptr(), s_pre.data_ptr(), idx.data_ptr()
    with torch.cuda.device(dev):
        _cabi.check(_cabi.pfpn_head_rsample_fwd(C.byref(a), _stream_ptr()))
    return sample, s_pre, idx


def rsample_bwd(logits, loc, logstd, g_sample, g_s_pre, *, seed=0, offset=0, ext_uniform=None, ext_normal=None,
                offset_dev=None):
    """Straight-through backward (mask / mask2, utils.py:164-183).  Returns (dlogits, dloc, dlogstd)."""
    a, keep = _rsample_args(logits, loc, logstd, seed, offset, ext_uniform, ext_normal, offset_dev)
    dev = logits.device
    g_sample = _f32c(g_sample, "g_sample")
    a.g_sample = g_sample.data_ptr()
    if g_s_pre is not None:
        g_s_pre = _f32c(g_s_pre, "g_s_pre")
        a.g_s_pre = g_s_pre.data_ptr()
    dlogits = torch.empty(a.B, a.A, a.P, dtype=torch.float32, device=dev)
    dloc = torch.zeros(a.A, a.P, dtype=torch.float32, device=dev)
    dlogstd = torch.zeros_like(dloc)
    a.dlogits, a.dloc, a.dlogstd = dlogits.data_ptr(), dloc.data_ptr(), dlogstd.data_ptr()
    n = C.c_size_t(0)
    _cabi.check(_cabi.pfpn_rsample_bwd_workspace_bytes(a.B, a.A, a.P, C.byref(n)))
    key = (dev, "rs", n.value)
    ws = _stats_ws.get(key)
    if ws is None:
        ws = _stats_ws[key] = torch.empty(n.value, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.pfpn_head_rsample_bwd(C.byref(a), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return dlogits, dloc, dlogstd


def rollout_fused(logits, loc, logstd, *, seed=0, offset=0, ext_uniform=None, ext_normal=None, max_active=None,
                  sum_active=None, want_ent=False):
    """K2f: sample + log_prob + entropy + activity statistics in one pass over the logits (utils.py:187-194,108-151,
    a2c.py:346-365).  Returns dict(action, idx, lp[, ent]); max_active / sum_active are updated in place when given.
    A = 36, P = 35 only (PfpnError -3 otherwise: the caller falls back to the three-kernel form)."""
    logits, loc, logstd = _f32c(logits, "logits"), _f32c(loc, "loc"), _f32c(logstd, "logstd")
    B, A, P = logits.shape
    dev = logits.device
    a = _cabi.RolloutArgs()
    a.logits, a.loc, a.logstd = logits.data_ptr(), loc.data_ptr(), logstd.data_ptr()
    keep = [logits, loc, logstd]
    if ext_uniform is not None:
        if ext_uniform.dtype != torch.float64 or ext_uniform.shape != (B, A):
            raise ValueError("ext_uniform must be float64 [B, A]")
        ext_uniform = ext_uniform.contiguous()
        keep.append(ext_uniform)
        a.ext_uniform = ext_uniform.data_ptr()
    if ext_normal is not None:
        ext_normal = _f32c(ext_normal, "ext_normal")
        keep.append(ext_normal)
        a.ext_normal = ext_normal.data_ptr()
    out = dict(action=torch.empty(B, A, dtype=torch.float32, device=dev), idx=_i32(B, A, device=dev),
               lp=torch.empty(B, dtype=torch.float32, device=dev))
    a.action, a.idx, a.lp = out["action"].data_ptr(), out["idx"].data_ptr(), out["lp"].data_ptr()
    if want_ent:
        out["ent"] = torch.empty(B, dtype=torch.float32, device=dev)
        a.ent = out["ent"].data_ptr()
    if (max_active is None) != (sum_active is None):
        raise ValueError("pass both max_active and sum_active or neither")
    ws_ptr, ws_n = None, 0
    if max_active is not None:
        for t, n in ((max_active, "max_active"), (sum_active, "sum_active")):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == (A, P)):
                raise ValueError(f"{n}: expected contiguous float32 CUDA [A, P]")
        a.max_active, a.sum_active = max_active.data_ptr(), sum_active.data_ptr()
        n = C.c_size_t(0)
        _cabi.check(_cabi.pfpn_rollout_workspace_bytes(A, P, C.byref(n)))
        key = (dev, "ro", n.value)
        ws = _stats_ws.get(key)
        if ws is None:
            ws = _stats_ws[key] = torch.empty(n.value, dtype=torch.uint8, device=dev)
        ws_ptr, ws_n = ws.data_ptr(), ws.numel()
    a.seed, a.offset, a.B, a.A, a.P = seed, offset, B, A, P
    with torch.cuda.device(dev):
        _cabi.check(_cabi.pfpn_head_rollout(C.byref(a), ws_ptr, ws_n, _stream_ptr()))
    return out


def sac_head_fused(logits, loc, logstd, g_sample, g_lp, *, seed=0, offset=0, ext_uniform=None, ext_normal=None,
                   out: Optional[dict] = None, dlogits_out=None, offset_dev=None):
    """K3f: rsample forward + tanh log_prob forward + the backward of both in one pass (utils.py:108-144,156-186).
    Returns dict(sample, s_pre, idx, logp, dlogits, dloc, dlogstd).  A = 36, P = 100 only (PfpnError -3 otherwise)."""
    logits, loc, logstd = _f32c(logits, "logits"), _f32c(loc, "loc"), _f32c(logstd, "logstd")
    g_sample, g_lp = _f32c(g_sample, "g_sample"), _f32c(g_lp, "g_lp")
    B, A, P = logits.shape
    if g_sample.shape != (B, A) or g_lp.shape != (B,):
        raise ValueError("g_sample [B, A], g_lp [B]")
    dev = logits.device
    a = _cabi.SacHeadArgs()
    a.logits, a.loc, a.logstd, a.g_sample, a.g_lp = (t.data_ptr() for t in (logits, loc, logstd, g_sample, g_lp))
    keep = [logits, loc, logstd, g_sample, g_lp]
    if (ext_uniform is None) != (ext_normal is None):
        raise ValueError("pass both ext_uniform and ext_normal or neither")
    if ext_uniform is not None:
        ext_uniform, ext_normal = _f32c(ext_uniform, "ext_uniform"), _f32c(ext_normal, "ext_normal")
        keep += [ext_uniform, ext_normal]
        a.ext_uniform, a.ext_normal = ext_uniform.data_ptr(), ext_normal.data_ptr()
    out = {} if out is None else out
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    if out.get("sample") is None:
        out.update(sample=f(B, A), s_pre=f(B, A), idx=_i32(B, A, device=dev), logp=f(B), dloc=f(A, P), dlogstd=f(A, P))
    if dlogits_out is not None:
        out["dlogits"] = dlogits_out
    if out.get("dlogits") is None:
        out["dlogits"] = f(B, A, P)
    a.sample, a.s_pre, a.idx, a.logp = (out[k].data_ptr() for k in ("sample", "s_pre", "idx", "logp"))
    a.dlogits, a.dloc, a.dlogstd = (out[k].data_ptr() for k in ("dlogits", "dloc", "dlogstd"))
    a.seed, a.offset, a.B, a.A, a.P = seed, offset, B, A, P
    if offset_dev is not None:
        assert offset_dev.dtype == torch.int64 and offset_dev.is_cuda
        a.offset_dev = offset_dev.data_ptr()
        keep.append(offset_dev)
    n = C.c_size_t(0)
    _cabi.check(_cabi.pfpn_sac_head_workspace_bytes(A, P, C.byref(n)))
    key = (dev, "sacf", n.value)
    ws = _stats_ws.get(key)
    if ws is None:
        ws = _stats_ws[key] = torch.empty(n.value, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.pfpn_sac_head_fwd_bwd(C.byref(a), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return out


def mean_action(logits, loc, tanh: bool = False):
    """utils.py:202-236 (forward).  Returns (action [B,A], idx [B,A])."""
    logits, loc = _f32c(logits, "logits"), _f32c(loc, "loc")
    B, A, P = logits.shape
    action = torch.empty(B, A, dtype=torch.float32, device=logits.device)
    idx = _i32(B, A, device=logits.device)
    with torch.cuda.device(logits.device):
        _cabi.check(_cabi.pfpn_head_mean(logits.data_ptr(), loc.data_ptr(), action.data_ptr(), idx.data_ptr(), B, A, P,
                                         _cabi.HEAD_FLAG_TANH if tanh else 0, _stream_ptr()))
    return action, idx


def stats_update(logits, max_active, sum_active, want_probs: bool = False):
    """a2c.py:356-360, in place on max_active / sum_active.  Returns probs [B,A,P] or None."""
    logits = _f32c(logits, "logits")
    B, A, P = logits.shape
    for t, n in ((max_active, "max_active"), (sum_active, "sum_active")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == (A, P)):
            raise ValueError(f"{n}: expected contiguous float32 CUDA [A, P]")
    probs = torch.empty_like(logits) if want_probs else None
    n = C.c_size_t(0)
    _cabi.check(_cabi.pfpn_stats_workspace_bytes(B, A, P, C.byref(n)))
    key = (logits.device, n.value)
    ws = _stats_ws.get(key)
    if ws is None:
        ws = _stats_ws[key] = torch.empty(n.value, dtype=torch.uint8, device=logits.device)
    with torch.cuda.device(logits.device):
        _cabi.check(_cabi.pfpn_stats_update(logits.data_ptr(), None if probs is None else probs.data_ptr(),
                                            max_active.data_ptr(), sum_active.data_ptr(), B, A, P, ws.data_ptr(), ws.numel(),
                                            _stream_ptr()))
    return probs


_stats_ws: dict = {}
