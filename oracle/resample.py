"""Index-level restatement of ``ParticleFilteringA2CNetwork.build_resample_ops``
(/root/reference/networks/actor_critic/a2c.py:385-474) and of the running activity
statistics (a2c.py:346-365).  TEST INFRASTRUCTURE ONLY.

All randomness is supplied by the caller:
  * ``cat_u``    fp64 [A, n] in [0,1)  -- draws of ``tf.random.categorical`` (a2c.py:402)
  * ``choice``   int32 [>=M]           -- draws of ``tf.random.uniform([M], 0, k, int32)`` (a2c.py:408)
  * ``noise_u``  fp32 [>=M] in [-1,1)  -- draws of ``tf.random.uniform(shape, -1, 1)`` (a2c.py:441)

Arithmetic conventions where the lowered graph does not pin the reference (un-vendored TF
kernels, "parity unpinned"), chosen so that CPU and GPU can agree bit-for-bit on indices:
  * row sums of ``sum_active`` are taken sequentially (k = 0..P-1) in fp32;
  * ``log(avg)`` is the fp64 log rounded to fp32 (TF: fp32 Eigen log);
  * the categorical draw follows TF-1.14's CPU ``Multinomial`` functor (fp64 CDF +
    upper_bound), see ``oracle.head.tf_multinomial_cpu``;
  * ``top_k(sorted=False)`` (resample > 0, used by no shipped setting) is returned in
    descending order, ties to the lower index.
"""
from __future__ import annotations

import numpy as np

from .head import tf_multinomial_cpu

F32 = np.float32


def activity_stats(probs: np.ndarray, max_active: np.ndarray, sum_active: np.ndarray):
    """a2c.py:356-360: running max / sum over the batch axis of softmax(logits) [B,A,P]."""
    probs = np.asarray(probs, dtype=F32)
    new_max = np.maximum(max_active.astype(F32), probs.max(axis=0))
    new_sum = sum_active.astype(F32) + probs.sum(axis=0, dtype=F32)
    return new_max, new_sum


def seq_rowsum_f32(x: np.ndarray) -> np.ndarray:
    out = np.zeros(x.shape[0], dtype=F32)
    for k in range(x.shape[1]):
        out = (out + x[:, k]).astype(F32)
    return out


def resample(max_active, sum_active, loc, logstd, bias, weight, *, resample=-1, threshold=None,
             tanh=False, cat_u=None, choice=None, noise_u=None):
    """Returns (new tensors dict, index dict).  Inputs are not modified."""
    max_active = np.asarray(max_active, dtype=F32)
    sum_active = np.asarray(sum_active, dtype=F32)
    loc = np.array(loc, dtype=F32)
    logstd = np.array(logstd, dtype=F32)
    bias = np.array(bias, dtype=F32)
    weight = np.array(weight, dtype=F32)
    A, n = max_active.shape
    assert weight.shape[1] == A * n and bias.shape == (A * n,)
    thr = F32(threshold if threshold else .05 / n)  # a2c.py:391

    # a2c.py:394-398
    rowsum = seq_rowsum_f32(sum_active)
    avg = (sum_active / rowsum[:, None]).astype(F32)
    invalid = np.argwhere(max_active < thr).astype(np.int32)  # row-major (a_m, j_m)
    M = invalid.shape[0]
    a_m, j_m = invalid[:, 0], invalid[:, 1]

    # a2c.py:400-408
    if resample < 0:
        assert resample == -1
        with np.errstate(divide="ignore"):
            logits = np.log(avg.astype(np.float64)).astype(F32)
        cand = tf_multinomial_cpu(logits, np.asarray(cat_u, dtype=np.float64)[:, :n])
        ch = j_m.copy()
        k = n
    else:
        k = min(n, resample)
        order = np.argsort(-avg, axis=1, kind="stable")  # descending, ties -> lower index
        cand = order[:, :k].astype(np.int32)
        ch = np.asarray(choice, dtype=np.int32)[:M]
    # a2c.py:410-413
    src = cand[a_m, ch].astype(np.int32) if M else np.zeros(0, np.int32)
    col = (a_m * n + j_m).astype(np.int32)
    tcol = (a_m * n + src).astype(np.int32)

    # a2c.py:420-427: gathers of PRE-update values
    std = np.exp(logstd).astype(F32)
    tloc = loc[a_m, src].astype(F32)
    tstd = std[a_m, src].astype(F32)
    tlogstd = logstd[a_m, src].astype(F32)
    tb = bias[tcol].astype(F32)
    tW = weight[:, tcol].copy()

    # a2c.py:441-445
    u = np.asarray(noise_u, dtype=F32)[:M] if M else np.zeros(0, F32)
    noise = (tstd * u).astype(F32)
    noise = (noise + np.where(noise < 0, F32(-1e-4), F32(1e-4))).astype(F32)
    tloc = (tloc + noise).astype(F32)
    # a2c.py:448-450 (fires for SAC: normalize_policy_output_ is True)
    if tanh:
        eps = F32(1e-6)
        tloc = np.arctanh(np.clip(tloc, eps - 1, 1 - eps).astype(F32)).astype(F32)
    tlogstd = np.clip(tlogstd, F32(-20), F32(2)).astype(F32)  # a2c.py:451

    # a2c.py:453-458
    uniq, first_pos, idx, count = np.unique(tcol, return_index=True, return_inverse=True, return_counts=True)
    # np.unique sorts; TF's unique_with_counts keeps first-occurrence order
    order = np.argsort(first_pos, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    uniq, count = uniq[order].astype(np.int32), count[order].astype(np.int32)
    idx = rank[idx].astype(np.int32) if M else np.zeros(0, np.int32)
    delta = np.array([np.sum(col == x) for x in uniq], dtype=F32)
    if M:
        tb = (tb - np.log((count.astype(F32) + F32(1) - delta).astype(F32))[idx]).astype(F32)

    # a2c.py:460-471
    loc[a_m, j_m] = tloc
    logstd[a_m, j_m] = tlogstd
    bias[tcol] = tb
    bias[col] = tb
    weight[:, col] = tW
    out = dict(loc=loc, logstd=logstd, bias=bias, weight=weight,
               max_active=np.zeros_like(max_active), sum_active=np.zeros_like(sum_active))  # a2c.py:372-378
    ints = dict(M=M, invalid=invalid, cand=cand, src=src, col=col, tcol=tcol, uniq=uniq, idx=idx,
                count=count, delta=delta.astype(np.int32))
    return out, ints


def mixture_weights(bias, weight, h):
    """softmax(h @ W + b) reshaped [B, A, P] -- used by the mass-conservation invariant."""
    z = h.astype(np.float64) @ weight.astype(np.float64) + bias.astype(np.float64)
    return z
