"""Op-order restatement of the reference's ``MixtureGaussianDistribution``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
``/root/reference/networks/utils.py:85-236`` one TF graph node per torch op, so
that an fp32 run rounds like the reference graph and an fp64 run is the
arbiter.  Gradients come from torch autograd through the same op chain, with
the reference's three ``tf.custom_gradient`` functions (``foo``
utils.py:109-117, ``mask2`` :164-171, ``mask`` :176-183, and the ``mean``
variant :213-220) re-created as ``torch.autograd.Function``s.

Random draws are always supplied by the caller (uniforms / normals), which is
what makes the CUDA kernels' verification mode comparable.
"""
from __future__ import annotations

import math

import numpy as np
import torch

HALF_LOG_2PI = 0.9189385175704956  # [graph] Normal/prob_1 const; == 0.5*ln(2*pi)
F32_TINY = float(np.finfo(np.float32).tiny)  # TFP 0.7 uniform minval


# --------------------------------------------------------------------------
# custom gradients
# --------------------------------------------------------------------------
class _Foo(torch.autograd.Function):
    """utils.py:109-117 -- identity whose incoming grad is zeroed where NaN/Inf."""

    @staticmethod
    def forward(ctx, p):
        return p.clone()

    @staticmethod
    def backward(ctx, dy):
        bad = torch.logical_or(torch.isnan(dy), torch.isinf(dy))
        return torch.where(bad, torch.zeros_like(dy), dy)


class _Mask2(torch.autograd.Function):
    """utils.py:164-171.  ``tanh_p`` and ``m`` are closed over (no grad path)."""

    @staticmethod
    def forward(ctx, w, p, m, tanh_p):
        y = m * p
        tanh_t = torch.sum(m * tanh_p, dim=-1, keepdim=True)
        ctx.save_for_backward(m, tanh_p, tanh_t)
        return y

    @staticmethod
    def backward(ctx, dy):
        m, tanh_p, tanh_t = ctx.saved_tensors
        gap = (tanh_p - tanh_t) / torch.clamp(1 - tanh_t ** 2, min=1e-6)
        return gap * dy, m * dy, None, None


class _Mask(torch.autograd.Function):
    """utils.py:176-183.  ``p`` here is ``tanh_p`` when normalize_output."""

    @staticmethod
    def forward(ctx, w, p, m):
        y = m * p
        t = torch.sum(y, dim=-1, keepdim=True)
        ctx.save_for_backward(m, p, t)
        return y

    @staticmethod
    def backward(ctx, dy):
        m, p, t = ctx.saved_tensors
        gap = p - t
        return gap * dy, m * dy, None


class _MeanMask(torch.autograd.Function):
    """utils.py:213-220 (deterministic action, tanh variant)."""

    @staticmethod
    def forward(ctx, w, p, m):
        y = m * p
        t = torch.sum(y, dim=-1, keepdim=True)
        ctx.save_for_backward(m, p, t)
        return y

    @staticmethod
    def backward(ctx, dy):
        m, p, t = ctx.saved_tensors
        gap = p - t
        return gap * dy, torch.sum(m * dy, dim=0, keepdim=True), None


# --------------------------------------------------------------------------
# third-party kernels restated
# --------------------------------------------------------------------------
def tf_softmax(logits: torch.Tensor) -> torch.Tensor:
    """TF ``Softmax`` kernel: exp(l - max) / sum (Eigen SoftmaxEigenImpl)."""
    shifted = logits - torch.amax(logits, dim=-1, keepdim=True)
    e = torch.exp(shifted)
    return e / torch.sum(e, dim=-1, keepdim=True)


def tf_log_softmax(logits: torch.Tensor) -> torch.Tensor:
    shifted = logits - torch.amax(logits, dim=-1, keepdim=True)
    return shifted - torch.log(torch.sum(torch.exp(shifted), dim=-1, keepdim=True))


def tf_normal_prob(x, loc, scale):
    """[graph] ``Normal/prob_1/*``: Sub, RealDiv, Square, Mul(-0.5), Log, Add, Sub, Exp."""
    z = (x - loc) / scale
    log_unnormalized = -0.5 * torch.square(z)
    log_normalization = HALF_LOG_2PI + torch.log(scale)
    return torch.exp(log_unnormalized - log_normalization)


def tf_multinomial_cpu(logits: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """TF-1.14 ``Multinomial`` CPU functor (core/kernels/multinomial_op.cc).

    ``logits`` [rows, classes] float32, ``uniforms`` [rows, num_samples] float64
    in [0,1).  Per row: max over finite logits; running fp64 CDF of
    ``exp(double(logit) - max)`` (non-finite logits add nothing); draw
    ``u * total``; index = ``upper_bound(cdf, u*total)``.  int32 output.
    The kernel is not vendored in /root/reference ("parity unpinned").
    """
    logits = np.asarray(logits, dtype=np.float32)
    uniforms = np.asarray(uniforms, dtype=np.float64)
    rows, classes = logits.shape
    out = np.empty(uniforms.shape, dtype=np.int32)
    for r in range(rows):
        row = logits[r]
        finite = np.isfinite(row)
        mx = np.float64(row[finite].max()) if finite.any() else np.float64(np.finfo(np.float32).min)
        e = np.where(finite, np.exp(row.astype(np.float64) - mx), 0.0)
        cdf = np.cumsum(e)  # sequential fp64 running total
        total = cdf[-1]
        # (u * total can round up to total; TF would then emit the out-of-range class `classes`,
        #  the oracle and the kernels clamp to the last class)
        out[r] = np.minimum(np.searchsorted(cdf, uniforms[r] * total, side="right"), classes - 1).astype(np.int32)
    return out


# --------------------------------------------------------------------------
# the distribution
# --------------------------------------------------------------------------
class MixtureGaussianOracle:
    """Same constructor / methods as the reference class (utils.py:85-236)."""

    def __init__(self, logits, loc, scale, normalize_output):
        self.logits, self.loc, self.scale = logits, loc, scale
        self.normalize_output = normalize_output
        self.probs = tf_softmax(logits)  # dis_dist.probs  (utils.py:97)
        self.dis_action = None

    # utils.py:108-134
    def log_prob(self, value):
        if self.normalize_output:
            if isinstance(value, (tuple, list)):
                value, value_before_tanh = value
            else:
                value_before_tanh = torch.atanh(value)
        else:
            value_before_tanh = value
        p = tf_normal_prob(value_before_tanh.unsqueeze(-1), self.loc, self.scale)
        p = torch.sum(self.probs * p, dim=-1)
        p = _Foo.apply(p)
        lp = torch.log(p)
        if self.normalize_output:
            lp = lp - 2 * (math.log(2.0) - value_before_tanh
                           - torch.nn.functional.softplus(-2 * value_before_tanh))
        return torch.sum(lp, dim=-1)

    def prob(self, value):  # utils.py:103-106
        return torch.exp(self.log_prob(value))

    # utils.py:146-151 -- categorical entropy per action dim, [B, A]
    def entropy(self):
        v = self.logits - torch.amax(self.logits, dim=-1, keepdim=True)
        s0 = torch.exp(v)
        s1 = torch.sum(s0, dim=-1, keepdim=True)
        p = s0 / s1
        return torch.sum(p * (torch.log(s1) - v), dim=-1)

    # utils.py:153-200
    def sample(self, n, *, uniform=None, normal=None):
        """``uniform``: plain branch fp64 [B,A] (Multinomial draw); rsample branch
        float [B,A,P] in [tiny,1).  ``normal``: standard normals [B,A,P]."""
        assert n == 1
        B, A, P = self.logits.shape
        if self.normalize_output:  # rsample, utils.py:156-186
            g = -torch.log(-torch.log(uniform))
            w = torch.exp(tf_log_softmax((g + self.logits) / 1.0))
            p = normal * self.scale + self.loc  # Normal.sample(B): [B,A,P]
            self.dis_action = torch.argmax(w, dim=-1)
            m = torch.nn.functional.one_hot(self.dis_action, P).to(w.dtype)
            tanh_p = torch.tanh(p)
            s_ = _Mask2.apply(w, p, m, tanh_p.detach())
            s_ = torch.sum(s_, -1).reshape(n, B, A)
            sample = _Mask.apply(w, tanh_p, m)
            sample = torch.sum(sample, -1).reshape(n, B, A)
            return sample, s_
        # plain branch, utils.py:187-194
        idx = tf_multinomial_cpu(
            self.logits.detach().to(torch.float32).reshape(B * A, P).numpy(),
            np.asarray(uniform, dtype=np.float64).reshape(B * A, 1))
        self.dis_action = torch.from_numpy(idx.reshape(B, A).astype(np.int64))
        p = normal * self.scale + self.loc
        mask = torch.nn.functional.one_hot(self.dis_action, P).to(p.dtype)
        sample = torch.sum(mask * p, -1)
        return sample.reshape(n, B, A)

    # utils.py:202-236
    def mean(self):
        B, A, P = self.logits.shape
        if self.normalize_output:
            w = tf_softmax(self.logits / 1.0)
            p = self.loc.unsqueeze(0)
            self.dis_action = torch.argmax(w, dim=-1)
            m = torch.nn.functional.one_hot(self.dis_action, P).to(w.dtype)
            p = torch.tanh(p)
            return torch.sum(_MeanMask.apply(w, p.expand(B, A, P), m), -1)
        dis_action = torch.argmax(self.logits, dim=-1)  # [B,A]
        return torch.gather(self.loc.unsqueeze(0).expand(B, A, P), 2,
                            dis_action.unsqueeze(-1)).squeeze(-1)


# --------------------------------------------------------------------------
# particle grid initialisation, a2c.py:476-535
# --------------------------------------------------------------------------
def init_particles(A: int, P: int, tanh: bool, init_sigma=None):
    """Returns (loc [A,P], logstd [A,P]) float64 numpy, bounds forced to +-1
    (a2c.py:479-480)."""
    u = np.ones((A, P))
    l = -np.ones((A, P))
    n = P
    if tanh:
        loc = l + (u - l) / n * (np.arange(n)[None, :] + 0.5)
    else:
        loc = l + (u - l) / (n - 1) * np.arange(n)[None, :]
    if init_sigma:
        std = np.full((A, P), float(init_sigma))
        if tanh:
            loc_ = loc
            loc = np.arctanh(loc)
            std = np.maximum(
                loc - np.arctanh(np.maximum(1e-6 - 1, loc_ - std)),
                np.arctanh(np.minimum(1 - 1e-6, loc_ + std)) - loc)
    else:
        std = (u - l) / (n - 1)
        if tanh:
            assert n > 3
            loc = np.arctanh(loc)
            std = np.empty_like(loc)
            for i in range(A):
                for j in range(P):
                    d0 = loc[i, j] - loc[i, max(0, j - 1)]
                    d1 = loc[i, min(n - 1, j + 1)] - loc[i, j]
                    std[i, j] = max(d0, d1)
    return loc, np.log(std)


# --------------------------------------------------------------------------
# head-facing losses
# --------------------------------------------------------------------------
def normalize_advantage(adv):
    """actor_critic.py:151-155: (adv - mean) / (sqrt(population var) + 1e-8)."""
    mean = torch.mean(adv)
    var = torch.mean(torch.square(adv - mean))
    return (adv - mean) / (torch.sqrt(var) + 1e-8)


def ppo_policy_loss(lp, lp_old, adv_n, eps=0.2):
    """ppo.py:44-54."""
    ratio = torch.exp(lp - lp_old)
    surrogate = ratio * adv_n
    clipped = torch.clamp(ratio, 1.0 - eps, 1.0 + eps) * adv_n
    return -torch.mean(torch.minimum(surrogate, clipped))


def sac_policy_loss(logp, q_min, log_alpha, A):
    """sac.py:166-173: mean(alpha*logp - min(Q1,Q2) - log_alpha * sg(logp + target_entropy)),
    alpha = sg(exp(log_alpha)) (sac.py:38), target_entropy = -A (sac.py:170)."""
    alpha = torch.exp(log_alpha).detach()
    l = alpha * logp - q_min
    l = l - log_alpha * (logp + (-A)).detach()
    return torch.mean(l)


def ppo_head_fwd_bwd(logits, loc, logstd, action, adv, lp_old, *, eps=0.2,
                     entropy_beta=0.0, normalize_adv=True, tanh=False, dtype=torch.float64):
    """One fused-K1 worth of reference work: log_prob, entropy, PPO surrogate
    loss and its gradients w.r.t. logits / loc / logstd (a2, a4, a13, A1)."""
    lg = logits.detach().to(dtype).requires_grad_(True)
    lc = loc.detach().to(dtype).requires_grad_(True)
    ls = logstd.detach().to(dtype).requires_grad_(True)
    dist = MixtureGaussianOracle(lg, lc, torch.exp(ls), tanh)
    act = action.to(dtype)
    lp = dist.log_prob((torch.tanh(act), act) if tanh else act)
    ent = dist.entropy()
    a = adv.to(dtype)
    a_n = normalize_advantage(a) if normalize_adv else a
    loss = ppo_policy_loss(lp, lp_old.to(dtype), a_n.detach(), eps)
    if entropy_beta:
        loss = loss - entropy_beta * torch.mean(torch.sum(ent, dim=1))
    loss.backward()
    return dict(lp=lp.detach(), ent=ent.detach(), loss=loss.detach(),
                dlogits=lg.grad, dloc=lc.grad, dlogstd=ls.grad)


def head_fwd_bwd(logits, loc, logstd, value, g_lp, g_ent=None, *, tanh=False,
                 dtype=torch.float64, want_dvalue=False):
    """Generic vjp: L = sum_b g_lp[b]*lp[b] + sum_{b,a} g_ent[b]*H[b,a]."""
    lg = logits.detach().to(dtype).requires_grad_(True)
    lc = loc.detach().to(dtype).requires_grad_(True)
    ls = logstd.detach().to(dtype).requires_grad_(True)
    v = value.detach().to(dtype).requires_grad_(want_dvalue)
    dist = MixtureGaussianOracle(lg, lc, torch.exp(ls), tanh)
    lp = dist.log_prob((torch.tanh(v), v) if tanh else v)
    ent = dist.entropy()
    L = torch.sum(g_lp.to(dtype) * lp)
    if g_ent is not None:
        L = L + torch.sum(g_ent.to(dtype).unsqueeze(-1) * ent)
    L.backward()
    out = dict(lp=lp.detach(), ent=ent.detach(), dlogits=lg.grad, dloc=lc.grad,
               dlogstd=ls.grad)
    if want_dvalue:
        out["dvalue"] = v.grad
    return out
