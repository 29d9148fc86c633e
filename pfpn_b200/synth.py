"""Seeded synthetic inputs of the shapes named in BASELINE.json.

Generated on the CPU with a seeded ``torch.Generator`` and then copied to the GPU,
so the CPU oracle and the CUDA kernels see identical bits.
"""
from __future__ import annotations

import math

import torch


def particle_grid(A: int, P: int, g: torch.Generator, jitter: bool = True):
    """Plain (non-tanh) particle init of a2c.py:476-500 plus a small jitter."""
    loc = torch.linspace(-1.0, 1.0, P).repeat(A, 1)
    logstd = torch.full((A, P), math.log(2.0 / (P - 1)))
    if jitter:
        loc = loc + 0.02 * torch.randn(A, P, generator=g)
        logstd = logstd + 0.1 * torch.randn(A, P, generator=g)
    return loc.contiguous(), logstd.contiguous()


def head_inputs(B: int, A: int, P: int, seed: int = 34114, far_frac: float = 0.01,
                logit_std: float = 2.0):
    """c2/c4-style head inputs: logits ~ N(0, 2^2); actions drawn from the mixture,
    a `far_frac` slice replaced by U(-3,3) outliers; adv ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, A, P, generator=g) * logit_std
    loc, logstd = particle_grid(A, P, g)
    idx = torch.multinomial(torch.softmax(logits.reshape(B * A, P), -1), 1, generator=g).reshape(B, A)
    mu = torch.gather(loc.expand(B, A, P), 2, idx[..., None])[..., 0]
    sd = torch.exp(torch.gather(logstd.expand(B, A, P), 2, idx[..., None])[..., 0])
    value = mu + sd * torch.randn(B, A, generator=g)
    n_far = int(far_frac * B)
    if n_far:
        value[:n_far] = torch.rand(n_far, A, generator=g) * 6 - 3
    adv = torch.randn(B, generator=g)
    lp_noise = 0.05 * torch.randn(B, generator=g)
    return dict(logits=logits, loc=loc, logstd=logstd, value=value.contiguous(), adv=adv, lp_noise=lp_noise)
