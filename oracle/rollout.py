"""ORACLE (test infrastructure, not product code): rollout post-processing of the A2C/PPO workers.

Restates, in numpy and in the reference's own operation order,
  * `discount`                          /root/reference/networks/utils.py:5-15
  * `generalized_advantage_estimate`    /root/reference/networks/actor_critic/a2c.py:30-40
  * `value_target_estimate`             /root/reference/networks/actor_critic/a2c.py:42-49
The reference scans a float32 array with Python-float factors.  Executed here (numpy 2.x, NEP 50) every step
is one float32 multiply and one float32 add -- that is what the golden file pins and what the kernel
reproduces bit for bit.  Under the numpy 1.x the reference was written for, scalar (float32, Python float)
arithmetic promotes to float64, i.e. the scan itself ran in float64 over the same float32 td errors; the two
differ by at most ~T * 2^-24 relative, far inside the 1e-5 tolerance (tests check both).
Pinned against the reference's own functions executed from source: tests/golden/gae_golden.npz
(oracle/gen_golden_gae.py).
"""
import numpy as np


def discount(val, factor, bootstrap_val):
    """utils.py:5-10 (normalize=False branch): result[t] = val[t] + factor * result[t+1], from the back."""
    val = np.asarray(val, dtype=np.float32)
    out = np.empty_like(val)
    run = np.float32(bootstrap_val)
    f = np.float32(factor)
    for t in range(len(val) - 1, -1, -1):
        run = np.float32(val[t] + np.float32(f * run))
        out[t] = run
    return out


def generalized_advantage_estimate(reward, value, gamma, gae_gamma):
    """a2c.py:30-40 (normalize=False): td = r + gamma*v' - v in float32, then `discount(td, gae_gamma, 0)`."""
    assert len(value) == len(reward) + 1
    r = np.asarray(reward, dtype=np.float32)
    v = np.asarray(value[:-1], dtype=np.float32)
    v_ = np.asarray(value[1:], dtype=np.float32)
    td = (r + np.float32(gamma) * v_ - v).astype(np.float32)  # (r + gamma*v_) - v, each op rounded to float32
    if gae_gamma:
        return discount(td, gae_gamma, 0.0)
    return td


def value_target_estimate(value, advantage):
    """a2c.py:42-49 (normalize=False): value + advantage."""
    return np.add(np.asarray(value, dtype=np.float32), np.asarray(advantage, dtype=np.float32))
