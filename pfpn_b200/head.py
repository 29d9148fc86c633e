"""Functional wrappers over the K1 C-ABI call (``pfpn_head_logprob``).

torch tensors are only the device-memory carrier: every function below hands raw
``data_ptr()``s and the current CUDA stream to ``libpfpn_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _cabi

_workspaces: dict = {}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ws(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (pfpn_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name}: expected float32, got {t.dtype}")
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def head_workspace_bytes(A: int, P: int) -> int:
    n = C.c_size_t(0)
    _cabi.check(_cabi.pfpn_head_workspace_bytes(A, P, C.byref(n)))
    return n.value


def launch_info(A: int, P: int, mode: int):
    out = (C.c_int32 * 4)()
    _cabi.check(_cabi.pfpn_head_launch_info(A, P, mode, out))
    return dict(num_sms=out[0], ctas_per_sm=out[1], threads=out[2], states_per_tile=out[3])


def adv_stats(adv: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """{mean, 1/(sqrt(popvar)+1e-8)} of the minibatch advantage (actor_critic.py:151-155)."""
    adv = _f32c(adv, "adv")
    if out is None:
        out = torch.empty(2, dtype=torch.float32, device=adv.device)
    _cabi.check(_cabi.pfpn_adv_stats(adv.data_ptr(), adv.numel(), out.data_ptr(), _stream_ptr()))
    return out


def head_call(mode: int, logits, loc, logstd, value, *, tanh=False, g_lp=None, g_ent_ba=None,
              g_ent: float = 0.0, adv=None, lp_old=None, adv_stats_t=None, eps_clip: float = 0.2,
              loss_scale: float = 0.0, want_ent_ba=False, want_dvalue=False, dlogits_out=None,
              out: Optional[dict] = None, push=None, consume_into=None) -> dict:
    """One launch of K1.  Returns a dict of freshly written tensors.

    ``out`` may carry preallocated ``lp, ent, dlogits, dloc, dlogstd, loss``
    tensors (bench / CUDA-graph use) -- then nothing is allocated here.
    ``push`` (a ``pfpn_b200.peer.PeerGather``): data-parallel form -- the finalize kernel also stores dloc / dlogstd
    into every rank's gather buffer; follow with ``push.reduce(...)``, or pass ``consume_into`` [2*A*P] and the NEXT
    call's launch writes the sum over ranks of this call's exchange there (one exchange late, no kernel of its own).
    """
    logits = _f32c(logits, "logits")
    B, A, P = logits.shape
    loc = _f32c(loc, "loc")
    logstd = _f32c(logstd, "logstd")
    value = _f32c(value, "value")
    if loc.shape != (A, P) or logstd.shape != (A, P) or value.shape != (B, A):
        raise ValueError("shape mismatch: logits[B,A,P], loc/logstd[A,P], value[B,A]")
    dev = logits.device
    out = {} if out is None else out
    new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    a = _cabi.HeadArgs()
    a.logits, a.loc, a.logstd, a.value = logits.data_ptr(), loc.data_ptr(), logstd.data_ptr(), value.data_ptr()
    a.B, a.A, a.P, a.mode = B, A, P, mode
    a.flags = _cabi.HEAD_FLAG_TANH if tanh else 0
    out.setdefault("lp", None)
    if out["lp"] is None:
        out["lp"] = new(B)
    if out.get("ent") is None:
        out["ent"] = new(B)
    a.lp, a.ent = out["lp"].data_ptr(), out["ent"].data_ptr()
    if want_ent_ba:
        if out.get("ent_ba") is None:
            out["ent_ba"] = new(B, A)
        a.ent_ba = out["ent_ba"].data_ptr()
    keep = [logits, loc, logstd, value]
    if mode != _cabi.HEAD_FWD:
        if dlogits_out is not None:
            out["dlogits"] = dlogits_out
        if out.get("dlogits") is None:
            out["dlogits"] = new(B, A, P)
        if out.get("dloc") is None:
            out["dloc"] = new(A, P)
        if out.get("dlogstd") is None:
            out["dlogstd"] = new(A, P)
        a.dlogits, a.dloc, a.dlogstd = out["dlogits"].data_ptr(), out["dloc"].data_ptr(), out["dlogstd"].data_ptr()
        a.g_ent = float(g_ent)
        if g_ent_ba is not None:
            g_ent_ba = _f32c(g_ent_ba, "g_ent_ba")
            keep.append(g_ent_ba)
            a.g_ent_ba = g_ent_ba.data_ptr()
        if want_dvalue:
            if out.get("dvalue") is None:
                out["dvalue"] = new(B, A)
            a.dvalue = out["dvalue"].data_ptr()
    if mode == _cabi.HEAD_GRAD:
        g_lp = _f32c(g_lp, "g_lp")
        keep.append(g_lp)
        a.g_lp = g_lp.data_ptr()
    elif mode == _cabi.HEAD_PPO:
        adv = _f32c(adv, "adv")
        lp_old = _f32c(lp_old, "lp_old")
        keep += [adv, lp_old]
        a.adv, a.lp_old = adv.data_ptr(), lp_old.data_ptr()
        if adv_stats_t is not None:
            keep.append(adv_stats_t)
            a.adv_stats = adv_stats_t.data_ptr()
        a.eps_clip = float(eps_clip)
        a.loss_scale = float(loss_scale if loss_scale else 1.0 / B)
        if out.get("loss") is None:
            out["loss"] = new(1)
        a.loss = out["loss"].data_ptr()
    nbytes = head_workspace_bytes(A, P)
    ws = _ws(dev, nbytes)
    with torch.cuda.device(dev):
        if push is not None and mode != _cabi.HEAD_FWD:
            _cabi.check(_cabi.pfpn_head_logprob_push(C.byref(a), ws.data_ptr(), ws.numel(), C.byref(push.push_args(consume_into)),
                                                     _stream_ptr()))
        else:
            _cabi.check(_cabi.pfpn_head_logprob(C.byref(a), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return out
