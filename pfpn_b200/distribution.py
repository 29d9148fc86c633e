"""Drop-in for the reference ``MixtureGaussianDistribution``
(/root/reference/networks/utils.py:85-236) backed by the sm_100a kernels.

Same constructor, method names, argument meaning and shape conventions; TF
symbolic tensors become fp32 CUDA ``torch.Tensor``s and TF autodiff becomes
``torch.autograd.Function``s whose ``backward`` launches the backward kernels.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi
from . import head as _head


class _LogProbFn(torch.autograd.Function):
    """log_prob [B] with the guarded gradient of utils.py:109-117."""

    @staticmethod
    def forward(ctx, logits, loc, logstd, value, tanh):
        out = _head.head_call(_cabi.HEAD_FWD, logits, loc, logstd, value, tanh=tanh)
        ctx.save_for_backward(logits, loc, logstd, value)
        ctx.tanh = tanh
        return out["lp"]

    @staticmethod
    def backward(ctx, g_lp):
        logits, loc, logstd, value = ctx.saved_tensors
        out = _head.head_call(_cabi.HEAD_GRAD, logits, loc, logstd, value, tanh=ctx.tanh,
                              g_lp=g_lp.contiguous(), want_dvalue=ctx.needs_input_grad[3])
        return (out["dlogits"], out["dloc"], out["dlogstd"],
                out.get("dvalue") if ctx.needs_input_grad[3] else None, None)


class _EntropyFn(torch.autograd.Function):
    """Categorical entropy per action dim [B, A] (utils.py:146-151)."""

    @staticmethod
    def forward(ctx, logits, loc, logstd):
        B, A, _ = logits.shape
        zeros = torch.zeros(B, A, dtype=torch.float32, device=logits.device)
        out = _head.head_call(_cabi.HEAD_FWD, logits, loc, logstd, zeros, want_ent_ba=True)
        ctx.save_for_backward(logits, loc, logstd, zeros)
        return out["ent_ba"]

    @staticmethod
    def backward(ctx, g_ent_ba):
        logits, loc, logstd, zeros = ctx.saved_tensors
        B = logits.shape[0]
        g0 = torch.zeros(B, dtype=torch.float32, device=logits.device)
        out = _head.head_call(_cabi.HEAD_GRAD, logits, loc, logstd, zeros, g_lp=g0,
                              g_ent_ba=g_ent_ba.contiguous())
        return out["dlogits"], None, None


class _DisDist:
    """Stand-in for ``dis_dist`` (utils.py:96-98): callers read ``.probs`` / ``.logits``."""

    def __init__(self, logits):
        self.logits = logits
        self._probs = None

    @property
    def probs(self):
        if self._probs is None:
            # consumed only by the running activity statistics (a2c.py:348-360),
            # for which `pfpn_b200.stats` has a fused kernel; this materialised
            # form exists for API parity.
            self._probs = torch.softmax(self.logits, dim=-1)
        return self._probs


class MixtureGaussianDistribution:
    def __init__(self, logits, loc, scale, normalize_output, *, logstd: Optional[torch.Tensor] = None):
        if logits.dim() != 3:
            raise NotImplementedError("only identical particle counts per dimension (utils.py:100-101)")
        self.logits, self.loc, self.scale = logits, loc, scale
        self.normalize_output = bool(normalize_output)
        # the kernels take log-std; the reference builds scale = exp(samples_std) (a2c.py:558)
        self.logstd = logstd if logstd is not None else torch.log(scale)
        self.dis_dist = _DisDist(logits)
        self.dis_action = None

    # utils.py:108-144
    def log_prob(self, value, name="log_prob"):
        if self.normalize_output:
            if isinstance(value, (tuple, list)):
                value, value_before_tanh = value
            else:
                value_before_tanh = torch.atanh(value)
        else:
            value_before_tanh = value
        return _LogProbFn.apply(self.logits, self.loc, self.logstd, value_before_tanh,
                                self.normalize_output)

    # utils.py:103-106
    def prob(self, value, name="prob"):
        return torch.exp(self.log_prob(value))

    # utils.py:146-151
    def entropy(self, name="entropy"):
        return _EntropyFn.apply(self.logits, self.loc, self.logstd)
