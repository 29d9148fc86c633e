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
from . import sampling as _sampling


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


class _LogProbEntropyFn(torch.autograd.Function):
    """log_prob [B] AND the categorical entropy [B, A] from ONE K1 forward; their gradients arrive together in ONE K1
    backward (dL/dlog_prob + dL/dentropy).  `MixtureGaussianDistribution.log_prob` goes through this node and keeps the
    entropy for a later `.entropy()` call on the same object, so a loss that uses both (A3C / IMPALA settings:
    policy loss + entropy_beta * entropy) costs two passes over [B,A,P] instead of four."""

    @staticmethod
    def forward(ctx, logits, loc, logstd, value, tanh):
        out = _head.head_call(_cabi.HEAD_FWD, logits, loc, logstd, value, tanh=tanh, want_ent_ba=True)
        ctx.save_for_backward(logits, loc, logstd, value)
        ctx.tanh = tanh
        return out["lp"], out["ent_ba"]

    @staticmethod
    def backward(ctx, g_lp, g_ent_ba):
        logits, loc, logstd, value = ctx.saved_tensors
        B = logits.shape[0]
        if g_lp is None:
            g_lp = torch.zeros(B, dtype=torch.float32, device=logits.device)
        out = _head.head_call(_cabi.HEAD_GRAD, logits, loc, logstd, value, tanh=ctx.tanh, g_lp=g_lp.contiguous(),
                              g_ent_ba=None if g_ent_ba is None else g_ent_ba.contiguous(),
                              want_dvalue=ctx.needs_input_grad[3])
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


class _RSampleFn(torch.autograd.Function):
    """(sample, s_) of utils.py:156-186 with the mask / mask2 custom gradients."""

    @staticmethod
    def forward(ctx, logits, loc, logstd, seed, offset, ext_uniform, ext_normal, offset_dev=None):
        sample, s_pre, idx = _sampling.rsample_fwd(logits, loc, logstd, seed=seed, offset=offset,
                                                   ext_uniform=ext_uniform, ext_normal=ext_normal, offset_dev=offset_dev)
        ctx.save_for_backward(logits, loc, logstd)
        ctx.rng = (seed, offset, ext_uniform, ext_normal, offset_dev)
        ctx.mark_non_differentiable(idx)
        return sample, s_pre, idx

    @staticmethod
    def backward(ctx, g_sample, g_s_pre, _g_idx):
        logits, loc, logstd = ctx.saved_tensors
        seed, offset, eu, en, odev = ctx.rng
        if g_sample is None:
            g_sample = torch.zeros(logits.shape[:2], dtype=torch.float32, device=logits.device)
        dlogits, dloc, dlogstd = _sampling.rsample_bwd(logits, loc, logstd, g_sample.contiguous(),
                                                       None if g_s_pre is None else g_s_pre.contiguous(),
                                                       seed=seed, offset=offset, ext_uniform=eu, ext_normal=en,
                                                       offset_dev=odev)  # (the word must not advance in between)
        return dlogits, dloc, dlogstd, None, None, None, None, None


class _DisDist:
    """Stand-in for ``dis_dist`` (utils.py:96-98): callers read ``.probs`` / ``.logits``."""

    def __init__(self, logits):
        self.logits = logits
        self._probs = None

    @property
    def probs(self):
        if self._probs is None:
            # the only consumer is the running activity statistics (a2c.py:348-360); the K4 kernel
            # produces the probabilities as a by-product (statistics go to scratch here).
            _, A, P = self.logits.shape
            scratch = torch.zeros(2, A, P, dtype=torch.float32, device=self.logits.device)
            self._probs = _sampling.stats_update(self.logits.detach(), scratch[0], scratch[1], want_probs=True)
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
        lp, ent_ba = _LogProbEntropyFn.apply(self.logits, self.loc, self.logstd, value_before_tanh, self.normalize_output)
        self._ent_ba = ent_ba  # same K1 forward; `.entropy()` on this object reuses it (one joint backward)
        return lp

    # utils.py:103-106
    def prob(self, value, name="prob"):
        return torch.exp(self.log_prob(value))

    # utils.py:146-151
    def entropy(self, name="entropy"):
        ent = getattr(self, "_ent_ba", None)
        if ent is not None:
            return ent
        return _EntropyFn.apply(self.logits, self.loc, self.logstd)

    # utils.py:153-200
    def sample(self, n, *, seed: int = 0, offset: int = 0, ext_uniform=None, ext_normal=None, offset_dev=None):
        """``n`` must be 1 (utils.py:154).  Plain branch -> [1, B, A]; rsample branch
        (normalize_output) -> tuple (sample [1,B,A], value_before_tanh [1,B,A]).
        Draws come from Philox(seed, offset [+ the int64 device word ``offset_dev``]) unless ext_* arrays are supplied."""
        assert n == 1
        B, A, _ = self.logits.shape
        if self.normalize_output:
            sample, s_pre, idx = _RSampleFn.apply(self.logits, self.loc, self.logstd, seed, offset, ext_uniform,
                                                  ext_normal, offset_dev)
            self.dis_action = idx
            return sample.reshape(n, B, A), s_pre.reshape(n, B, A)
        action, idx = _sampling.sample_plain(self.logits.detach(), self.loc.detach(), self.logstd.detach(), seed=seed,
                                             offset=offset, ext_uniform=ext_uniform, ext_normal=ext_normal)
        self.dis_action = idx
        return action.reshape(n, B, A)

    # utils.py:202-236 (forward; the reference only evaluates it on non-trainable evaluator nets)
    def mean(self):
        if not hasattr(self, "_determinstic_action"):
            action, idx = _sampling.mean_action(self.logits.detach(), self.loc.detach(), tanh=self.normalize_output)
            self.dis_action = idx
            self._determinstic_action = action
        return self._determinstic_action
