"""Drop-in for the reference's ``ParticleFilteringClipPPONetwork`` on the learner-update path
(/root/reference/networks/actor_critic/{actor_critic,a2c,ppo}.py), eager instead of a TF graph.

Same constructor keywords, ``init()``, ``run`` / ``evaluate`` / ``train`` signatures and return
conventions (``sess`` / ``ops`` are ignorable handles; ``optimizer`` is a
``pfpn_b200.learner.SyncReplicasAdam``).  All parameters live in ONE flat fp32 buffer in the
reference's variable order, all gradients in one flat bucket that also carries the pushed
statistics, so the data-parallel exchange is a single all-reduce.

Compute = hand-written CUDA behind the C ABI (K1 head, K2 sampling, K4 statistics, K5 resampling,
K6 trunk GEMMs); torch is the allocator and the stream provider.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from . import head as _head
from . import resampling as _resampling
from . import sampling as _sampling
from .head import _stream_ptr


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class _Linear:
    """fc_layer (ops.py:82-118): W [in, out] row-major + b, views into the flat buffers."""

    def __init__(self, name, k_in, n_out, k_pad=None):
        self.name, self.k_in, self.n_out = name, k_in, n_out
        self.k = k_pad or k_in  # rows of the stored W (zero rows beyond k_in)
        self.W = self.b = self.dW = self.db = None

    def numel(self):
        return self.k * self.n_out + self.n_out


def initial_particles(A: int, P: int, tanh: bool):
    """a2c.py:476-535 with the bounds forced to +-1 (:479-480): (loc [A,P], logstd [A,P]) fp32 on the host.
    Plain: linspace(-1, 1, P), std = 2/(P-1).  tanh (normalize_policy_output_): cell-centred grid mapped through
    arctanh, std = the larger neighbour gap (:486,501-512)."""
    if tanh:
        assert P > 3
        c = -1.0 + (2.0 / P) * (np.arange(P) + 0.5)
        mu = np.arctanh(c)
        sd = np.array([max(mu[j] - mu[max(0, j - 1)], mu[min(P - 1, j + 1)] - mu[j]) for j in range(P)])
    else:
        mu = -1.0 + 2.0 / (P - 1) * np.arange(P)
        sd = np.full(P, 2.0 / (P - 1))
    return (torch.tensor(mu, dtype=torch.float32).repeat(A, 1), torch.tensor(np.log(sd), dtype=torch.float32).repeat(A, 1))


class ParticleFilteringClipPPONetwork:
    GLOBAL_STEP0 = 0

    def __init__(self, trainable, state_shape, action_shape, action_lower_bound=None, action_upper_bound=None,
                 init_sigma=None, fixed_sigma=False, particles=50, resample=3, resample_interval=2000,
                 resample_threshold=None, normalize_policy_output=False, epsilon=0.2,
                 common_net_shape=(), critic_net_shape=(1024, 512), actor_net_shape=(1024, 512),
                 weight_initializer=None, activator="relu6", normalize_state=False, clip_state=False,
                 normalize_value=False, clip_value=False, normalize_advantage=False, clip_advantage=False,
                 critic_regularizer=None, actor_regularizer=None, entropy_beta=None, value_loss_coef=0.5,
                 gamma=0.99, lambd=0.95, log=True, device="cuda", seed=0, **kwargs):
        if common_net_shape:
            raise NotImplementedError("common_net_shape is [] in every shipped setting (deepmimic_base.py:4-6)")
        if fixed_sigma or init_sigma:
            raise NotImplementedError("fixed_sigma / init_sigma are not used by the shipped PFPN settings")
        if normalize_value or clip_value or clip_advantage or critic_regularizer or actor_regularizer:
            raise NotImplementedError("value normaliser / regularisers are off in every shipped setting")
        if activator not in ("relu6",) and getattr(activator, "__name__", "") != "relu6":
            raise NotImplementedError("only relu6 (deepmimic_base.py:8)")
        self.trainable = bool(trainable)
        self.random_action = self.trainable
        self.state_shape = list(state_shape) if hasattr(state_shape, "__len__") else [state_shape]
        self.action_shape = list(action_shape) if hasattr(action_shape, "__len__") else [action_shape]
        if len(self.action_shape) != 1:
            raise ValueError("Particle Filtering Policy Network only supports continuous action space.")
        self.S, self.A, self.P = int(self.state_shape[0]), int(self.action_shape[0]), int(particles)
        self.dis_action_shape = [self.P] * self.A
        self.action_lower_bound, self.action_upper_bound = action_lower_bound, action_upper_bound
        self.resample, self.resample_interval, self.resample_threshold = resample, resample_interval, resample_threshold
        self.normalize_policy_output = False              # a2c.py:327
        self.normalize_policy_output_ = bool(normalize_policy_output)  # a2c.py:328
        self.epsilon = epsilon
        self.actor_net_shape, self.critic_net_shape = list(actor_net_shape), list(critic_net_shape)
        self.normalize_state, self.clip_state = bool(normalize_state), float(clip_state or 0.0)
        self.normalize_advantage = bool(normalize_advantage)
        self.entropy_beta, self.value_loss_coef = entropy_beta, value_loss_coef
        self.gamma, self.gae_gamma = gamma, None if lambd is None else gamma * lambd
        self.device = torch.device(device)
        # `seed` is REPLICATED state: weight / particle initialisation and the resample tick's draws must be identical on
        # every data-parallel rank (there is no parameter broadcast).  The rollout / SAC sampling draws instead come from a
        # per-rank Philox stream (`sample_seed`, derived in init() from the rank) so shards do not share their noise.
        self.seed = int(seed)
        self.sample_seed = int(seed)
        self.init_ops, self.train_ops, self.running_update_ops = [], [], []
        self.local_update_variables: List[torch.Tensor] = []
        self.global_step = self.GLOBAL_STEP0
        self._rng_offset = 0
        self._act = {}
        # K6 path: tcgen05 3xTF32 GEMMs (within the fp32 tolerance, tests/test_tc_gemm_gpu.py) unless the
        # fp32 FFMA anchor is requested
        import os
        self.use_tensor_cores = os.environ.get("PFPN_TRUNK", "tc") != "ffma"
        self.use_presplit = os.environ.get("PFPN_PRESPLIT", "1") != "0"

    # ------------------------------------------------------------------------------ build ----
    def _derive_sample_stream(self):
        import torch.distributed as dist
        rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
        self.sample_seed = (self.seed ^ (0x9E3779B97F4A7C15 * rank)) & 0xFFFFFFFFFFFFFFFF

    def init(self):
        S, A, P, dev = self.S, self.A, self.P, self.device
        self._derive_sample_stream()
        self.Sp = _pad4(S)
        dims_a = [self.Sp] + self.actor_net_shape
        dims_c = [self.Sp] + self.critic_net_shape
        self.actor = [_Linear(f"actor/fc{i+1}", (S if i == 0 else dims_a[i]), dims_a[i + 1], dims_a[i])
                      for i in range(len(self.actor_net_shape))]
        self.fc_policy = _Linear("actor/fc_policy", dims_a[-1], A * P)
        self.critic = [_Linear(f"critic/fc{i+1}", (S if i == 0 else dims_c[i]), dims_c[i + 1], dims_c[i])
                       for i in range(len(self.critic_net_shape))]
        self.critic.append(_Linear(f"critic/fc{len(self.critic)+1}", dims_c[-1], 1))
        # flat layout in the reference's variable order ([graph] / SURVEY 2.2 C1)
        order = [("lin", l) for l in self.actor] + [("samples", None), ("samples_std", None), ("lin", self.fc_policy)] + \
                [("lin", l) for l in self.critic]
        self.n_stats = 2 * S + 2 * A * P
        off = 0

        def take(cnt, shape):
            nonlocal off
            assert off % 4 == 0, "every tensor starts 16-byte aligned"
            p, g = self.params[off:off + cnt].view(*shape), self.grads[off:off + cnt].view(*shape)
            off += _pad4(cnt)
            return p, g

        # (padding tensors to multiples of 4 floats keeps every view 16-byte aligned for the kernels)
        self.n_params = n = _pad4(sum(_pad4(l.k * l.n_out) + _pad4(l.n_out) if k == "lin" else _pad4(A * P)
                                      for k, l in order))
        self.params = torch.zeros(n, dtype=torch.float32, device=dev)
        # params - tf32(params): the low halves of the weights for the 3xTF32 GEMMs' B operand, split once per optimizer
        # step (written by the optimizer kernel) instead of once per tile per GEMM (K6)
        self.params_lo = torch.zeros(n, dtype=torch.float32, device=dev)
        self._lo_version = -1
        # gradient bucket = [grads | pushed statistics] : one all-reduce (sync_model.py:92-96)
        self.bucket = torch.zeros(n + _pad4(self.n_stats), dtype=torch.float32, device=dev)
        self.grads = self.bucket[:n]
        for kind, l in order:
            if kind == "lin":
                l.W_lo = self.params_lo[off:off + l.k * l.n_out].view(l.k, l.n_out)
                l.W, l.dW = take(l.k * l.n_out, (l.k, l.n_out))
                l.b, l.db = take(l.n_out, (l.n_out,))
            elif kind == "samples":
                self.loc, self.dloc = take(A * P, (A, P))
            else:
                self.logstd, self.dlogstd = take(A * P, (A, P))
        self.policy_weight, self.policy_bias = self.fc_policy.W, self.fc_policy.b  # a2c.py:546-551
        self._init_values()
        # non-trainable state (checkpointed / synchronised like the reference's local_update_variables)
        self.state_mean = torch.zeros(S, dtype=torch.float32, device=dev)
        self.state_std = torch.ones(S, dtype=torch.float32, device=dev)
        self.max_active = torch.zeros(A, P, dtype=torch.float32, device=dev)
        self.sum_active = torch.zeros(A, P, dtype=torch.float32, device=dev)
        self.train_flag = 0
        if self.normalize_state:
            self.local_update_variables += [self.state_mean, self.state_std]   # actor_critic.py:333
        if self.trainable and self.resample:
            self.local_update_variables += [self.max_active, self.sum_active]  # a2c.py:362-363
            self.train_ops.append(self.update)                                 # a2c.py:383
        self._scratch = torch.empty(max(2 * S, 8), dtype=torch.float32, device=dev)
        self._init_step_state()
        return self

    def _init_step_state(self):
        """Device-resident step counters {exchange calls, Adam step, global_step, ticket} (read by the kernels of the
        graph-capturable update, csrc/syncstep.cu) + their host shadow, buffers of the pushed statistics, side stream."""
        dev, S = self.device, self.S
        self.dev_counters = torch.zeros(4, dtype=torch.int32, device=dev)
        self._dev_shadow = [0, 0, 0]
        self._new_mean = torch.zeros(S, dtype=torch.float32, device=dev)
        self._new_std = torch.ones(S, dtype=torch.float32, device=dev)
        n = C.c_size_t(0)
        _cabi.check(_cabi.pfpn_normalizer_scratch_bytes(S, C.byref(n)))
        self._norm_scratch = torch.empty(n.value // 8, dtype=torch.float64, device=dev)
        self._adv_stats = torch.empty(2, dtype=torch.float32, device=dev)
        self._side = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        import os
        # the critic trunk is independent of the actor trunk + head: run it on a second stream (a parallel branch of the
        # captured graph); pays off when a rank's minibatch shard no longer fills the GPU with one GEMM
        self.overlap_critic = os.environ.get("PFPN_CRITIC_STREAM", "1") != "0"
        self.overlap_wgrad = os.environ.get("PFPN_WGRAD_STREAM", "1") != "0"

    def _ensure_counters(self, calls=None, adam_step=None):
        """Upload the host-side step numbers when the device copy is stale (first use, checkpoint resume, a switch between
        exchange implementations).  Steady state: no-op -- the kernels increment the device copy themselves."""
        want = [self._dev_shadow[0] if calls is None else int(calls),
                self._dev_shadow[1] if adam_step is None else int(adam_step), int(self.global_step)]
        if want != self._dev_shadow:
            self.dev_counters.copy_(torch.tensor(want + [0], dtype=torch.int32), non_blocking=False)
            self._dev_shadow = want

    def _init_values(self):
        """a2c.py:476-535 particle grid (bounds forced to +-1) + truncated_normal(0, .01) weights."""
        A, P = self.A, self.P
        g = torch.Generator().manual_seed(self.seed)
        loc, logstd = initial_particles(A, P, self.normalize_policy_output_)
        self.loc.copy_(loc)
        self.logstd.copy_(logstd)
        for l in self.actor + [self.fc_policy] + self.critic:
            w = torch.empty(l.k_in, l.n_out)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=0.01, a=-0.02, b=0.02, generator=g)
            l.W.zero_()
            l.W[:l.k_in].copy_(w)
            l.b.zero_()

    # ---------------------------------------------------------------------------- forward ----
    def _buf(self, name, *shape):
        t = self._act.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(*shape, dtype=torch.float32, device=self.device)
            self._act[name] = t
        return t

    def _lo_ok(self):
        """The low halves follow the parameters: the optimizer kernel writes them; any OTHER in-place change of
        ``self.params`` (initialisation, checkpoint load, a torch op -- they bump the tensor's version counter) or a kernel
        that edits parameters through raw pointers (resample tick, legacy optimizer path: they call ``invalidate_lo``)
        triggers one re-split here."""
        if not self.use_presplit:
            return False
        if self._lo_version != self.params._version:
            _cabi.check(_cabi.pfpn_split_lo(self.params.data_ptr(), self.params_lo.data_ptr(), self.n_params, _stream_ptr()))
            self._lo_version = self.params._version
        return True

    def invalidate_lo(self):
        self._lo_version = -1

    def _linear(self, l: _Linear, X, Y, relu6):
        if self.use_tensor_cores and l.n_out > 1:
            # W [in, out] as stored is the MN-major B operand: no transposed copy of the weights
            if self._lo_ok():
                _cabi.check(_cabi.pfpn_tc_gemm_nn_lo(X.data_ptr(), X.stride(0), l.W.data_ptr(), l.W_lo.data_ptr(), l.n_out,
                                                     Y.data_ptr(), Y.stride(0), l.b.data_ptr(), None, 0, X.shape[0], l.n_out,
                                                     l.k, 2 if relu6 else 1, _stream_ptr()))
                return
            _cabi.check(_cabi.pfpn_tc_gemm_nn(X.data_ptr(), X.stride(0), l.W.data_ptr(), l.n_out, Y.data_ptr(), Y.stride(0),
                                              l.b.data_ptr(), None, 0, X.shape[0], l.n_out, l.k, 2 if relu6 else 1,
                                              _stream_ptr()))
            return
        _cabi.check(_cabi.pfpn_mlp_linear_fwd(X.data_ptr(), X.stride(0), l.W.data_ptr(), l.b.data_ptr(), Y.data_ptr(),
                                              Y.stride(0) if Y.dim() > 1 else 1, X.shape[0], l.k, l.n_out,
                                              1 if relu6 else 0, _stream_ptr()))

    def _normalized(self, state: torch.Tensor):
        B = state.shape[0]
        x = self._buf("x", B, self.Sp)
        _cabi.check(_cabi.pfpn_state_normalize(state.data_ptr(), self.state_mean.data_ptr(), self.state_std.data_ptr(),
                                               x.data_ptr(), B, self.S, self.Sp, self.clip_state,
                                               1 if self.normalize_state else 0, _stream_ptr()))
        return x

    def _forward_actor(self, x):
        B, h, acts = x.shape[0], x, [x]
        for i, l in enumerate(self.actor):
            y = self._buf(f"h{i}", B, l.n_out)
            self._linear(l, h, y, True)
            h = y
            acts.append(y)
        logits = self._buf("logits", B, self.A * self.P)
        self._linear(self.fc_policy, h, logits, False)
        return logits.view(B, self.A, self.P), acts

    def _forward_critic(self, x):
        B, h, cacts = x.shape[0], x, [x]
        for i, l in enumerate(self.critic[:-1]):
            y = self._buf(f"c{i}", B, l.n_out)
            self._linear(l, h, y, True)
            h = y
            cacts.append(y)
        value = self._buf("value", B)
        self._linear(self.critic[-1], h, value, False)
        return value, cacts

    def _forward(self, state: torch.Tensor, want_value=True):
        x = self._normalized(state)
        logits, acts = self._forward_actor(x)
        value, cacts = self._forward_critic(x) if want_value else (None, [x])
        return logits, acts, value, cacts

    def _dev_state(self, state):
        t = torch.as_tensor(np.asarray(state, dtype=np.float32) if not torch.is_tensor(state) else state)
        t = t.to(self.device, dtype=torch.float32, non_blocking=True)
        return t.reshape(-1, self.S).contiguous()

    # ---------------------------------------------------------------------- rollout side ----
    def run_batch(self, state, ext_uniform=None, ext_normal=None):
        """Vectorised ``run``: (action [B,A], log_prob [B], value [B]); updates the activity
        statistics like the reference's running_update_ops do on every rollout step."""
        s = self._dev_state(state)
        logits, _, value, _ = self._forward(s)
        want_stats = self.trainable and self.resample
        if self.random_action and not self.normalize_policy_output_ and self.A == 36 and self.P == 35 and \
                os.environ.get("PFPN_ROLLOUT_FUSED", "1") != "0":
            # K2f: sample + log_prob + activity statistics in ONE pass over the logits (the shipped DPPO-PFPN shape)
            out = _sampling.rollout_fused(logits, self.loc, self.logstd, seed=self.sample_seed, offset=self._rng_offset,
                                          ext_uniform=ext_uniform, ext_normal=ext_normal,
                                          max_active=self.max_active if want_stats else None,
                                          sum_active=self.sum_active if want_stats else None)
            self._rng_offset += 2
            return out["action"], out["lp"], value.clone()
        if self.random_action:
            if self.normalize_policy_output_:
                smp, s_pre, _ = _sampling.rsample_fwd(logits, self.loc, self.logstd, seed=self.sample_seed, offset=self._rng_offset,
                                                      ext_uniform=ext_uniform, ext_normal=ext_normal)
                action, val = smp, s_pre
            else:
                action, _ = _sampling.sample_plain(logits, self.loc, self.logstd, seed=self.sample_seed, offset=self._rng_offset,
                                                   ext_uniform=ext_uniform, ext_normal=ext_normal)
                val = action
            self._rng_offset += 2
        else:
            action, _ = _sampling.mean_action(logits, self.loc, tanh=self.normalize_policy_output_)
            val = torch.atanh(action) if self.normalize_policy_output_ else action
        out = _head.head_call(_cabi.HEAD_FWD, logits, self.loc, self.logstd, val, tanh=self.normalize_policy_output_)
        if want_stats:
            _sampling.stats_update(logits, self.max_active, self.sum_active)
        return action, out["lp"], value.clone()

    def run(self, sess, state, ops=None):
        """ppo.py:56-62 + actor_critic.py:368-380: one state in, every output un-batched."""
        a, lp, v = self.run_batch(np.asarray(state, dtype=np.float32)[None])
        res = [a[0].cpu().numpy()]
        if self.trainable:
            res += [float(lp[0]), float(v[0])]
        return res

    def evaluate(self, sess, state):
        """a2c.py:75-78."""
        _, _, value, _ = self._forward(self._dev_state(np.asarray(state, dtype=np.float32)[None]))
        return float(value[0])

    # ------------------------------------------------------------- rollout post-processing ----
    def generalized_advantage_estimate(self, reward, value, normalize=False):
        """a2c.py:30-40 on the device, batched over trajectories: reward [E,T] (or [T]), value [E,T+1] (or [T+1],
        bootstrap value last) -> advantage with the same leading shape.  ``gae_gamma = gamma*lambd`` as in a2c.py:21."""
        adv, _ = self._gae(reward, value, want_target=False)
        if normalize:  # a2c.py:36-40 normalises over the whole trajectory (host-side option, unused by DPPO)
            adv = (adv - adv.mean()) / (adv.std(unbiased=False) + (1e-6 if self.gae_gamma else 1e-8))
        return adv

    def advantage_and_value_target(self, reward, value):
        """(advantage, value_target = value[:-1] + advantage): what the PPO worker feeds as `advantage` and
        what setup_value_target_tensor rebuilds as adv + v_old (workers/ppo.py:43-73, ppo.py:31-34)."""
        return self._gae(reward, value, want_target=True)

    def _gae(self, reward, value, want_target):
        r = torch.as_tensor(reward, dtype=torch.float32).to(self.device)
        v = torch.as_tensor(value, dtype=torch.float32).to(self.device)
        squeeze = r.dim() == 1
        r, v = r.reshape(-1, r.shape[-1]).contiguous(), v.reshape(-1, v.shape[-1]).contiguous()
        E, T = r.shape
        if v.shape != (E, T + 1):
            raise ValueError("value must hold one more entry per trajectory than reward (the bootstrap value)")
        adv = torch.empty_like(r)
        tgt = torch.empty_like(r) if want_target else None
        _cabi.check(_cabi.pfpn_gae(r.data_ptr(), v.data_ptr(), adv.data_ptr(), tgt.data_ptr() if want_target else None, E, T,
                                   float(self.gamma), float(self.gae_gamma or 0.0), _stream_ptr()))
        if squeeze:
            adv, tgt = adv[0], (tgt[0] if want_target else None)
        return adv, tgt

    # ------------------------------------------------------------------------ train step ----
    def compute_gradients(self, state, action, value, log_prob, advantage, loss_scale: Optional[float] = None):
        """Forward + backward of loss = policy_loss + value_loss_coef * value_loss on this rank's
        minibatch (ppo.py:39-54, actor_critic.py:128-184); gradients land in ``self.grads``.
        Returns device scalars (loss, entropy | None, policy_loss, value_loss)."""
        s = self._dev_state(state)
        B = s.shape[0]
        dv = lambda t: torch.as_tensor(np.asarray(t, dtype=np.float32) if not torch.is_tensor(t) else t).to(
            self.device, dtype=torch.float32).contiguous()
        action, value_old, lp_old, adv = dv(action).reshape(B, self.A), dv(value).reshape(B), dv(log_prob).reshape(B), \
            dv(advantage).reshape(B)
        scale = loss_scale if loss_scale else 1.0 / B
        # statistics pushed with this step (LocalUpdateHookPre, sync_model.py:123-138): computed from
        # the pre-update values, applied by the optimizer after aggregation
        if self.normalize_state:
            self._ensure_counters()
            _cabi.check(_cabi.pfpn_normalizer_update_dev(s.data_ptr(), self.state_mean.data_ptr(), self.state_std.data_ptr(),
                                                         self._new_mean.data_ptr(), self._new_std.data_ptr(), B, self.S,
                                                         self.dev_counters[2:].data_ptr(), self._norm_scratch.data_ptr(),
                                                         self._norm_scratch.numel() * 8, _stream_ptr()))
        x = self._normalized(s)
        vloss = self._buf("vloss", 1)

        def critic_branch():
            v, cacts = self._forward_critic(x)
            dval = self._buf("dval", B)
            _cabi.check(_cabi.pfpn_value_loss(v.data_ptr(), adv.data_ptr(), value_old.data_ptr(), dval.data_ptr(),
                                              vloss.data_ptr(), B, float(self.value_loss_coef), scale, _stream_ptr()))
            self._backward_stack(self.critic, cacts, dval, tag="c")

        main = torch.cuda.current_stream(self.device)
        fork = self.overlap_critic and self._side is not None
        if fork:  # parallel branch: critic trunk forward, value loss, critic backward
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                critic_branch()
        logits, acts = self._forward_actor(x)
        stats = _head.adv_stats(adv, out=self._adv_stats) if self.normalize_advantage else None
        ent_scale = -float(self.entropy_beta) * scale if self.entropy_beta else 0.0
        # action_hist is the stored (tanh'd) action: MixtureGaussianDistribution.log_prob applies atanh to a non-tuple
        # value when normalize_output (utils.py:120-126); K1 takes the pre-tanh value
        head_value = torch.atanh(action) if self.normalize_policy_output_ else action
        out = _head.head_call(_cabi.HEAD_PPO, logits, self.loc, self.logstd, head_value, tanh=self.normalize_policy_output_,
                              adv=adv, lp_old=lp_old, adv_stats_t=stats, eps_clip=self.epsilon, loss_scale=scale,
                              g_ent=ent_scale, dlogits_out=logits,
                              out=dict(dloc=self.dloc, dlogstd=self.dlogstd, lp=self._buf("lp", B), ent=self._buf("ent", B),
                                       loss=self._buf("ploss", 1)))
        dlogits = logits.view(B, self.A * self.P)
        self._backward_stack(self.actor + [self.fc_policy], acts, dlogits, tag="a")
        if fork:
            main.wait_stream(self._side)
        else:
            critic_branch()
        policy_loss = out["loss"][0]
        entropy = None
        if self.entropy_beta:
            entropy = out["ent"].sum() * scale
            policy_loss = policy_loss - self.entropy_beta * entropy
        value_loss = self.value_loss_coef * vloss[0]  # actor_critic.py:131-133: the returned value_loss is the weighted one
        loss = policy_loss + value_loss
        return loss, entropy, policy_loss, value_loss

    def _backward_stack(self, layers: Sequence[_Linear], acts, dY, tag=""):
        """acts[i] is the input of layers[i]; dY is dL/d(output of the last layer).  `tag` names the scratch buffer (the
        actor and critic stacks may run concurrently on two streams)."""
        st = _stream_ptr()
        # The weight gradient and the input gradient of a layer both only READ dY: the weight-gradient GEMMs run on an
        # auxiliary stream (a further parallel branch of the captured graph), so that their CTAs fill the SMs the
        # input-gradient GEMM's last wave leaves idle -- at 8192 states per GPU no GEMM of this net fills 148 SMs evenly.
        cur = torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
        aux = self._aux_stream(tag) if (getattr(self, "overlap_wgrad", False) and cur is not None) else None
        aux_used = False
        for i in range(len(layers) - 1, -1, -1):
            l, X = layers[i], acts[i]
            M = X.shape[0]
            if self.use_tensor_cores and l.n_out > 1 and M >= 512:
                if aux is not None:
                    aux.wait_stream(cur)  # dY of this layer (and everything before it) is complete on `cur`
                    aux_used = True
                    with torch.cuda.stream(aux):
                        self._tc_wgrad(l, X, dY, M, tag + "w")
                else:
                    self._tc_wgrad(l, X, dY, M, tag)
                if i > 0:
                    dX = self._buf(f"d_{l.name}", M, l.k)
                    if self._lo_ok():
                        _cabi.check(_cabi.pfpn_tc_gemm_nt_lo(dY.data_ptr(), dY.stride(0), l.W.data_ptr(), l.W_lo.data_ptr(), l.n_out,
                                                             dX.data_ptr(), dX.stride(0), None, X.data_ptr(), X.stride(0), M, l.k,
                                                             l.n_out, 3, st))
                    else:
                        _cabi.check(_cabi.pfpn_tc_gemm_nt(dY.data_ptr(), dY.stride(0), l.W.data_ptr(), l.n_out, dX.data_ptr(),
                                                          dX.stride(0), None, X.data_ptr(), X.stride(0), M, l.k, l.n_out, 3, st))
                    dY = dX
                continue
            n = C.c_size_t(0)
            _cabi.check(_cabi.pfpn_mlp_wgrad_workspace_bytes(M, l.k, l.n_out, C.byref(n)))
            ws = self._ws(n.value, tag)
            ldy = dY.stride(0) if dY.dim() > 1 else 1
            _cabi.check(_cabi.pfpn_mlp_linear_bwd_weight(X.data_ptr(), X.stride(0), dY.data_ptr(), ldy, l.dW.data_ptr(),
                                                         l.db.data_ptr(), M, l.k, l.n_out, ws.data_ptr(), ws.numel(), st))
            if i > 0:  # no gradient into the (stop_gradient) normalised state
                dX = self._buf(f"d_{l.name}", M, l.k)
                if self.use_tensor_cores and l.n_out > 1:
                    # dX[M, k] = (dY[M, n] W[k, n]^T) .* relu6'(X): W as stored is already the K-major operand
                    _cabi.check(_cabi.pfpn_tc_gemm_nt(dY.data_ptr(), ldy, l.W.data_ptr(), l.n_out, dX.data_ptr(), dX.stride(0),
                                                      None, X.data_ptr(), X.stride(0), M, l.k, l.n_out, 3, st))
                    dY = dX
                    continue
                _cabi.check(_cabi.pfpn_mlp_linear_bwd_input(dY.data_ptr(), ldy, l.W.data_ptr(), X.data_ptr(), dX.data_ptr(),
                                                            dX.stride(0), M, l.k, l.n_out, st))
                dY = dX
        if aux_used:  # (joining a stream that never forked would break a graph capture)
            cur.wait_stream(aux)

    def _aux_stream(self, tag):
        d = self.__dict__.setdefault("_aux_streams", {})
        if tag not in d:
            d[tag] = torch.cuda.Stream(self.device)
        return d[tag]

    def _tc_wgrad(self, l, X, dY, M, tag=""):
        """dW and db on the tensor cores: X and dY as stored (MN-major operands), split-K GEMM; the bias
        gradient is accumulated from the dY tiles the GEMM stages anyway."""
        st = _stream_ptr()
        n = C.c_size_t(0)
        _cabi.check(_cabi.pfpn_tc_wgrad_workspace_bytes(M, l.k, l.n_out, C.byref(n)))
        ws = self._ws(n.value, tag)
        _cabi.check(_cabi.pfpn_tc_linear_bwd_weight(X.data_ptr(), X.stride(0), dY.data_ptr(), dY.stride(0), l.dW.data_ptr(),
                                                    l.db.data_ptr(), M, l.k, l.n_out, ws.data_ptr(), ws.numel(), st))

    def _ws(self, nbytes, tag=""):
        t = self._act.get("_ws" + tag)
        if t is None or t.numel() < nbytes:
            t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._act["_ws" + tag] = t
        return t

    def train(self, sess, optimizer, ops, state, action, value, log_prob, advantage):
        """ppo.py:64-72 -> actor_critic.py:382-404: ((loss, entropy|None, policy_loss, value_loss), extra)."""
        scale = 1.0 / max(1, np.shape(advantage)[0])
        losses = self.compute_gradients(state, action, value, log_prob, advantage, loss_scale=scale)
        if optimizer is not None:
            optimizer.apply_gradients(self)
        extra = [] if not ops else [None] * (len(ops) if hasattr(ops, "__len__") else 1)
        return tuple(None if x is None else float(x) for x in losses), extra

    # ------------------------------------------------------------- resample tick (a2c.py:370-383) ----
    def update(self):
        self.train_flag += 1
        if self.train_flag >= self.resample_interval:
            _resampling.resample_(self.max_active, self.sum_active, self.loc, self.logstd, self.policy_bias,
                                  self.policy_weight, resample=self.resample, threshold=self.resample_threshold,
                                  tanh=self.normalize_policy_output_, seed=self.seed + 1, offset=3 * self.global_step)
            self.train_flag = 0
            self.invalidate_lo()  # the resampler rewired fc_policy columns through raw pointers:
            self._lo_ok()         # re-split now (a graph-replayed next step never runs the Python check)
            return True
        return False

    # -------------------------------------------------------------------- (de)serialisation ----
    def state_dict(self):
        return dict(params=self.params.clone(), state_mean=self.state_mean.clone(), state_std=self.state_std.clone(),
                    max_active=self.max_active.clone(), sum_active=self.sum_active.clone(), train_flag=self.train_flag,
                    global_step=self.global_step)

    def load_state_dict(self, sd):
        self.params.copy_(sd["params"])
        for k in ("state_mean", "state_std", "max_active", "sum_active"):
            getattr(self, k).copy_(sd[k])
        self.train_flag, self.global_step = int(sd["train_flag"]), int(sd["global_step"])

    def named_parameters(self):
        """Reference variable names ([.index] of the shipped checkpoint) -> (param view, grad view)."""
        out = {}
        for l in self.actor + [self.fc_policy] + self.critic:
            out[f"global_net/{l.name}/weight"] = (l.W[:l.k_in], l.dW[:l.k_in])
            out[f"global_net/{l.name}/bias"] = (l.b, l.db)
        out["global_net/actor/samples"] = (self.loc, self.dloc)
        out["global_net/actor/samples_std"] = (self.logstd, self.dlogstd)
        return out
