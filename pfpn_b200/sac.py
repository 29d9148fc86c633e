"""SAC-PFPN learner step: twin Q critics on [state | action], target critics with the
soft update, the learned temperature, the two-loss / two-optimizer update -- on top of the PFPN head kernels.

Mirrors `ParticleFilteringSACNetwork = sac_network_wrapper(ParticleFilteringA2CNetwork)`
(/root/reference/networks/actor_critic/sac.py:12-179) and the DDPG/SAC worker's optimizer
(models/workers/base_worker.py:25-120 with `separate_optimizer = True`, models/workers/ddpg.py:39-42):

  policy at s        a, s_ = policy.sample(1)  (K3: reparameterised, tanh), logp = policy.log_prob((a, s_))  (K1, tanh)
  critics            q{1,2}(s, a)  [gradient flows into a only]   and   q{1,2}(s, a_hist)
  target             vf' = min(q1', q2')(s', a') - alpha logp(a'|s') with a' from the ONLINE actor (shared template,
                     sac.py:147-149) and the target critics; q_target = r + gamma nt vf'
  losses             value_loss = coef mean((q_t - q1)^2 + (q_t - q2)^2); policy_loss = mean(alpha logp - min q - log_alpha
                     sg(logp - A))   (pfpn_sac_losses)
  update             c_grads over {critic, log_alpha(None)}, a_grads over {actor, log_alpha}; ONE joint global-norm clip
                     (lr_actor == lr_critic branch of clip_grads); critic Adam, then actor Adam; train_ops: resample tick,
                     soft target sync v_ <- (1 - tau) v_ + tau v.

The trunk / critic GEMMs, the head kernels and the optimizer kernels are the ones of the DPPO path; the only new
device code is `pfpn_sac_losses` and `pfpn_axpby`.  `torch.autograd` is used for the head composition only (the
autograd Functions of `distribution.py` call K1 / K3), exactly as a reference-side user would.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _cabi
from .distribution import MixtureGaussianDistribution
from .head import _stream_ptr
from .learner import SyncReplicasAdam, assert_replicas_identical
from .network import ParticleFilteringClipPPONetwork, _Linear, _pad4


class ParticleFilteringSACNetwork(ParticleFilteringClipPPONetwork):
    TRAIN_RNG_BASE = 1 << 40

    def __init__(self, trainable, state_shape, action_shape, alpha=0.2, tau=0.005, **kwargs):
        kwargs.setdefault("normalize_policy_output", True)  # sac.py:15-16
        kwargs.setdefault("gamma", 0.95)                    # deepmimic_base.py:12
        super().__init__(trainable, state_shape, action_shape, **kwargs)
        self.tau = float(tau)  # (the ctor's alpha is overwritten by exp(log_alpha), log_alpha = 0: sac.py:37-39)

    # ------------------------------------------------------------------------------ build ----
    def init(self):
        S, A, P, dev = self.S, self.A, self.P, self.device
        self._derive_sample_stream()
        self.use_presplit = False  # (the SAC optimizer keeps no pre-split weight halves: the GEMMs split on the fly)
        self.Sp, self.QI = _pad4(S), _pad4(S + A)
        dims_a = [self.Sp] + self.actor_net_shape
        self.actor = [_Linear(f"actor/fc{i+1}", (S if i == 0 else dims_a[i]), dims_a[i + 1], dims_a[i])
                      for i in range(len(self.actor_net_shape))]
        self.fc_policy = _Linear("actor/fc_policy", dims_a[-1], A * P)
        dims_c = [self.QI] + self.critic_net_shape

        def qnet(prefix):
            ls = [_Linear(f"{prefix}/fc{i+1}", (S + A if i == 0 else dims_c[i]), dims_c[i + 1], dims_c[i])
                  for i in range(len(self.critic_net_shape))]
            return ls + [_Linear(f"{prefix}/fc{len(ls)+1}", dims_c[-1], 1)]

        self.q = [qnet("critic/q1"), qnet("critic/q2")]
        self.qt = [qnet("target_net/critic/q1"), qnet("target_net/critic/q2")]
        self.critic = []  # (the PPO value head does not exist here: build_value returns None, sac.py:151-154)
        lin_numel = lambda l: _pad4(l.k * l.n_out) + _pad4(l.n_out)
        self.n_critic = sum(lin_numel(l) for net in self.q for l in net)
        n_actor = sum(lin_numel(l) for l in self.actor + [self.fc_policy]) + 2 * _pad4(A * P)
        self.n_params = n = self.n_critic + n_actor + 4  # [critic segment | actor segment | log_alpha (+3 pad)]
        self.n_stats = 2 * S + 2 * A * P
        self.params = torch.zeros(n, dtype=torch.float32, device=dev)
        self.bucket = torch.zeros(n + _pad4(self.n_stats), dtype=torch.float32, device=dev)
        self.grads = self.bucket[:n]
        self.target_params = torch.zeros(self.n_critic, dtype=torch.float32, device=dev)
        off = 0

        def take(buf, gbuf, cnt, shape):
            nonlocal off
            p = buf[off:off + cnt].view(*shape)
            g = None if gbuf is None else gbuf[off:off + cnt].view(*shape)
            off += _pad4(cnt)
            return p, g

        for net in self.q:
            for l in net:
                l.W, l.dW = take(self.params, self.grads, l.k * l.n_out, (l.k, l.n_out))
                l.b, l.db = take(self.params, self.grads, l.n_out, (l.n_out,))
        assert off == self.n_critic
        for l in self.actor:
            l.W, l.dW = take(self.params, self.grads, l.k * l.n_out, (l.k, l.n_out))
            l.b, l.db = take(self.params, self.grads, l.n_out, (l.n_out,))
        self.loc, self.dloc = take(self.params, self.grads, A * P, (A, P))
        self.logstd, self.dlogstd = take(self.params, self.grads, A * P, (A, P))
        l = self.fc_policy
        l.W, l.dW = take(self.params, self.grads, l.k * l.n_out, (l.k, l.n_out))
        l.b, l.db = take(self.params, self.grads, l.n_out, (l.n_out,))
        self.log_alpha, self.dlog_alpha = self.params[off:off + 1], self.grads[off:off + 1]
        assert off + 4 == n
        off = 0
        for net in self.qt:
            for l in net:
                l.W, _ = take(self.target_params, None, l.k * l.n_out, (l.k, l.n_out))
                l.b, _ = take(self.target_params, None, l.n_out, (l.n_out,))
        self.policy_weight, self.policy_bias = self.fc_policy.W, self.fc_policy.b
        self._init_values()  # particles + truncated-normal weights of self.actor + fc_policy (+ self.critic: empty)
        g = torch.Generator().manual_seed(self.seed + 17)
        for net in self.q:
            for l in net:
                w = torch.empty(l.k_in, l.n_out)
                torch.nn.init.trunc_normal_(w, mean=0.0, std=0.01, a=-0.02, b=0.02, generator=g)
                l.W.zero_()
                l.W[:l.k_in].copy_(w)
                l.b.zero_()
        self.target_params.copy_(self.params[:self.n_critic])  # init_target_net (sac.py:75-88)
        self.state_mean = torch.zeros(S, dtype=torch.float32, device=dev)
        self.state_std = torch.ones(S, dtype=torch.float32, device=dev)
        self.max_active = torch.zeros(A, P, dtype=torch.float32, device=dev)
        self.sum_active = torch.zeros(A, P, dtype=torch.float32, device=dev)
        self.train_flag = 0
        if self.normalize_state:
            self.local_update_variables += [self.state_mean, self.state_std]
        if self.trainable and self.resample:
            self.local_update_variables += [self.max_active, self.sum_active]
            self.train_ops.append(self.update)             # PFPN init runs inside super().init() (sac.py:41) ...
        if self.trainable:
            self.train_ops.append(self.sync_target_net)    # ... before the target sync is appended (sac.py:67-73)
        self._scratch = torch.empty(max(2 * S, 8), dtype=torch.float32, device=dev)
        # Device-resident step state: nothing the training step's kernels are launched with changes between steps, so the
        # whole step can be captured in a CUDA graph (GraphedSACUpdate).  `_train_rng` is added to the Philox offset of the
        # training draws by the kernels themselves and advanced by 4 at the end of every compute_gradients; `_gstep_dev`
        # mirrors global_step for the normaliser's decay (optim.cu).
        self._train_rng = torch.zeros(1, dtype=torch.int64, device=dev)
        self._gstep_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._new_mean, self._new_std = self.state_mean.clone(), self.state_std.clone()
        nb = C.c_size_t(0)
        _cabi.check(_cabi.pfpn_normalizer_scratch_bytes(S, C.byref(nb)))
        self._norm_scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        return self

    # ---------------------------------------------------------------------------- forward ----
    def _normalize(self, state, tag, out=None):
        B = state.shape[0]
        x = self._buf(f"x{tag}", B, self.Sp) if out is None else out
        _cabi.check(_cabi.pfpn_state_normalize(state.data_ptr(), self.state_mean.data_ptr(), self.state_std.data_ptr(),
                                               x.data_ptr(), B, self.S, self.Sp, self.clip_state,
                                               1 if self.normalize_state else 0, _stream_ptr()))
        return x

    def _actor_forward(self, x, tag):
        h, acts = x, [x]
        for i, l in enumerate(self.actor):
            y = self._buf(f"a{tag}h{i}", x.shape[0], l.n_out)
            self._linear(l, h, y, True)
            h = y
            acts.append(y)
        logits = self._buf(f"a{tag}logits", x.shape[0], self.A * self.P)
        self._linear(self.fc_policy, h, logits, False)
        return logits, acts

    def _q_forward(self, layers: List[_Linear], x, action, tag):
        """q(s, a) = build_value(concat([x, a])) (sac.py:107-113); returns (q [B], the layer inputs)."""
        B = x.shape[0]
        xin = self._buf(f"q{tag}in", B, self.QI)
        xin.zero_()
        xin[:, :self.S].copy_(x[:, :self.S])
        xin[:, self.S:self.S + self.A].copy_(action)
        h, acts = xin, [xin]
        for i, l in enumerate(layers[:-1]):
            y = self._buf(f"q{tag}h{i}", B, l.n_out)
            self._linear(l, h, y, True)
            h = y
            acts.append(y)
        q = self._buf(f"q{tag}out", B)
        self._linear(layers[-1], h, q, False)
        return q, acts

    def _input_grad(self, layers: List[_Linear], acts, dq, tag):
        """dL/d(input of the stack) for dL/dq = dq [B], no weight gradients (the policy loss skips the critic variables,
        sac.py:167).  Returns dL/dxin [B, QI]."""
        st = _stream_ptr()
        dY, ldy = dq, 1
        for i in range(len(layers) - 1, -1, -1):
            l, X = layers[i], acts[i]
            B = X.shape[0]
            dX = self._buf(f"dq{tag}_{i}", B, l.k)
            if self.use_tensor_cores and l.n_out > 1 and B >= 512:
                # dX = (dY W^T) [.* relu6'(X)]: W [k, n_out] as stored is the K-major operand of the tcgen05 GEMM
                _cabi.check(_cabi.pfpn_tc_gemm_nt(dY.data_ptr(), ldy, l.W.data_ptr(), l.n_out, dX.data_ptr(), dX.stride(0), None,
                                                  X.data_ptr() if i > 0 else None, X.stride(0), B, l.k, l.n_out,
                                                  3 if i > 0 else 0, st))
            else:
                _cabi.check(_cabi.pfpn_mlp_linear_bwd_input(dY.data_ptr(), ldy, l.W.data_ptr(), X.data_ptr() if i > 0 else None,
                                                            dX.data_ptr(), dX.stride(0), B, l.k, l.n_out, st))
            dY, ldy = dX, dX.stride(0)
        return dY

    def _parallel(self, thunks):
        """Run independent pieces of the step on side streams (fork after the current stream, join back into it).  At the
        reference's batch of 256 the step is a chain of ~150 kernels of 5-10 us each: the six critic evaluations, the two
        critic backward stacks and the two action-gradient chains have no data dependence on each other, so as parallel
        branches (of the eager streams and of the captured graph alike) they shorten the critical path."""
        import os
        if len(thunks) < 2 or self.device.type != "cuda" or os.environ.get("PFPN_SAC_STREAMS", "1") == "0":
            return [t() for t in thunks]
        if not torch.cuda.is_current_stream_capturing() and not getattr(self, "_wide", False):
            return [t() for t in thunks]  # (eager small batches are host-bound: the extra stream switches only cost time)
        cur = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_pstreams"):
            self._pstreams = [torch.cuda.Stream(self.device) for _ in range(6)]
            self._pdone = [torch.cuda.Event() for _ in range(6)]
            self._pfork = torch.cuda.Event()
        self._pfork.record(cur)
        outs = []
        for k, t in enumerate(thunks):
            s = self._pstreams[k % len(self._pstreams)]
            s.wait_event(self._pfork)
            with torch.cuda.stream(s):
                outs.append(t())
        for k in range(min(len(thunks), len(self._pstreams))):
            self._pdone[k].record(self._pstreams[k])
            cur.wait_event(self._pdone[k])
        return outs

    def _policy(self, logits, B, requires_grad, seed_offset, ext, offset_dev=None):
        lg = logits.view(B, self.A, self.P)
        loc, ls = self.loc, self.logstd
        if requires_grad:
            lg = lg.detach().requires_grad_(True)
            loc, ls = loc.detach().requires_grad_(True), ls.detach().requires_grad_(True)
        dist = MixtureGaussianDistribution(lg, loc, torch.exp(ls.detach()), True, logstd=ls)
        kw = dict(ext_uniform=ext[0], ext_normal=ext[1]) if ext is not None else \
            dict(seed=self.sample_seed, offset=seed_offset, offset_dev=offset_dev)
        smp, s_ = dist.sample(1, **kw)
        logp = dist.log_prob((smp[0], s_[0]))
        return smp[0], logp, (lg, loc, ls)

    # ---------------------------------------------------------------------- rollout side ----
    def run(self, sess, state, ops=None):
        """sac.py:91-95: the sampled action only."""
        a, _, _ = self.run_batch(np.asarray(state, dtype=np.float32)[None])
        return [a[0].cpu().numpy()] + ([None] * len(ops) if ops else [])

    def run_batch(self, state, ext_uniform=None, ext_normal=None):
        s = self._dev_state(state)
        logits, _ = self._actor_forward(self._normalize(s, "r"), "r")
        with torch.no_grad():
            a, logp, _ = self._policy(logits, s.shape[0], False, self._rng_offset,
                                      None if ext_uniform is None else (ext_uniform, ext_normal))
        self._rng_offset += 2
        if self.trainable and self.resample:
            from . import sampling as _sampling
            _sampling.stats_update(logits.view(s.shape[0], self.A, self.P), self.max_active, self.sum_active)
        return a, logp, None

    def evaluate(self, sess, state):
        raise NotImplementedError("SAC has no state-value head (sac.py:151-154)")

    # ------------------------------------------------------------------------ train step ----
    def compute_gradients(self, state, action, reward, not_terminal, state_, draws=None):
        """Forward + backward of value_loss (critic variables) and policy_loss (actor variables + log_alpha) on this
        rank's minibatch; gradients land in ``self.grads``.  ``draws`` = (U, EPS, U_next, EPS_next) reproduces given
        Gumbel uniforms / location normals (verification); None draws from Philox.
        Returns device scalars (loss, None, policy_loss, value_loss)."""
        dv = lambda t, *shape: torch.as_tensor(np.asarray(t, dtype=np.float32) if not torch.is_tensor(t) else t).to(
            self.device, dtype=torch.float32).reshape(*shape).contiguous()
        s, s2 = self._dev_state(state), self._dev_state(state_)
        B = s.shape[0]
        self._wide = B >= 2048  # (eager: side streams only where the kernels are long enough to overlap)
        a_hist, r, nt = dv(action, B, self.A), dv(reward, B), dv(not_terminal, B)
        st = _stream_ptr()
        if self.normalize_state:  # LocalUpdateHookPre: statistics of this minibatch, applied by the optimizer
            _cabi.check(_cabi.pfpn_normalizer_update_dev(s.data_ptr(), self.state_mean.data_ptr(), self.state_std.data_ptr(),
                                                         self._new_mean.data_ptr(), self._new_std.data_ptr(), B, self.S,
                                                         self._gstep_dev.data_ptr(), self._norm_scratch.data_ptr(),
                                                         self._norm_scratch.numel(), st))
        # the actor sees s and s' through the SAME weights: one forward over 2B rows (half the launches of this part of the
        # step, and from batch 256 on the stacked rows reach the tensor-core GEMM path)
        xx = self._buf("xx", 2 * B, self.Sp)
        x, x2 = self._normalize(s, "0", out=xx[:B]), self._normalize(s2, "1", out=xx[B:])
        logits_all, acts_all = self._actor_forward(xx, "01")
        logits, logits2 = logits_all[:B], logits_all[B:]
        a_acts = [t[:B] for t in acts_all]
        # training draws: Philox offsets TRAIN_RNG_BASE + {0 (s), 2 (s')} + the device word (rollouts count up from 0 on the host)
        off, odev = self.TRAIN_RNG_BASE, (self._train_rng if draws is None else None)
        # K3f (A = 36, P = 100): the head's backward -- rsample backward + tanh log_prob forward/backward -- is ONE pass that
        # regenerates the forward's draws; the forward sample below then needs no autograd graph
        import os
        fused = (self.A, self.P) == (36, 100) and os.environ.get("PFPN_SAC_FUSED", "1") != "0"
        smp, logp, leaves = self._policy(logits, B, not fused, off, None if draws is None else (draws[0], draws[1]), odev)
        with torch.no_grad():
            a2, logp2, _ = self._policy(logits2, B, False, off + 2, None if draws is None else (draws[2], draws[3]), odev)
        a_det = smp.detach()
        qs = self._parallel([lambda i=i: self._q_forward(self.q[i], x, a_det, f"a{i}") for i in range(2)] +
                            [lambda i=i: self._q_forward(self.q[i], x, a_hist, f"r{i}") for i in range(2)] +
                            [lambda i=i: self._q_forward(self.qt[i], x2, a2, f"t{i}") for i in range(2)])
        q_a, q_r, q_t = qs[0:2], qs[2:4], qs[4:6]
        f = lambda name: self._buf(name, B)
        dq_a, dq_r, dlogp, out4 = [f("dq1a"), f("dq2a")], [f("dq1r"), f("dq2r")], f("dlogp"), self._buf("sac_out", 4)
        _cabi.check(_cabi.pfpn_sac_losses(q_a[0][0].data_ptr(), q_a[1][0].data_ptr(), q_r[0][0].data_ptr(), q_r[1][0].data_ptr(),
                                          q_t[0][0].data_ptr(), q_t[1][0].data_ptr(), logp.detach().data_ptr(),
                                          logp2.data_ptr(), r.data_ptr(), nt.data_ptr(), self.log_alpha.data_ptr(),
                                          float(self.gamma), float(self.value_loss_coef), -float(self.A), B,
                                          dq_a[0].data_ptr(), dq_a[1].data_ptr(), dq_r[0].data_ptr(), dq_r[1].data_ptr(),
                                          dlogp.data_ptr(), out4.data_ptr(), st))
        self.grads.zero_()
        # value loss -> critic variables; policy loss: through the critics into the action (no critic weight gradients),
        # then the head, then the actor.  Four independent chains (disjoint gradient segments / scratch buffers).
        res = self._parallel([lambda i=i: self._backward_stack(self.q[i], q_r[i][1], dq_r[i], f"q{i}") for i in range(2)] +
                             [lambda i=i: self._input_grad(self.q[i], q_a[i][1], dq_a[i], f"a{i}") for i in range(2)])
        da = res[2][:, self.S:self.S + self.A] + res[3][:, self.S:self.S + self.A]
        if fused:
            from . import sampling as _sampling
            ext = {} if draws is None else dict(ext_uniform=draws[0], ext_normal=draws[1])
            fo = _sampling.sac_head_fused(logits.view(B, self.A, self.P), self.loc, self.logstd, da.contiguous(), dlogp,
                                          seed=self.sample_seed, offset=off, offset_dev=odev, **ext)
            self._backward_stack(self.actor + [self.fc_policy], a_acts, fo["dlogits"].view(B, self.A * self.P))
            self.dloc.copy_(fo["dloc"])
            self.dlogstd.copy_(fo["dlogstd"])
        else:
            lg, loc_l, ls_l = leaves
            torch.autograd.backward([logp, smp], [dlogp, da])
            self._backward_stack(self.actor + [self.fc_policy], a_acts, lg.grad.reshape(B, self.A * self.P).contiguous())
            self.dloc.copy_(loc_l.grad)
            self.dlogstd.copy_(ls_l.grad)
        self.dlog_alpha.copy_(out4[2:3])
        if draws is None:
            self._train_rng.add_(4)  # (after the backward: it regenerates the forward's draws from the same word)
        return out4[0] + out4[1], None, out4[1], out4[0]

    def train(self, sess, optimizer, ops, state, action, reward, not_terminal, state_):
        """sac.py:97-105 -> actor_critic.py:382-404: ((loss, None, policy_loss, value_loss), extra)."""
        losses = self.compute_gradients(state, action, reward, not_terminal, state_)
        if optimizer is not None:
            optimizer.apply_gradients(self)
        extra = [] if not ops else [None] * (len(ops) if hasattr(ops, "__len__") else 1)
        return tuple(None if v is None else float(v) for v in losses), extra

    def sync_target_net(self):
        """v_ <- (1 - tau) v_ + tau v over the critic variables (sac.py:67-73)."""
        _cabi.check(_cabi.pfpn_axpby(self.target_params.data_ptr(), self.params.data_ptr(), self.n_critic, 1.0 - self.tau,
                                     self.tau, _stream_ptr()))

    def named_parameters(self):
        out = {}
        for l in self.actor + [self.fc_policy] + [l for net in self.q for l in net]:
            out[f"global_net/{l.name}/weight"] = (l.W[:l.k_in], l.dW[:l.k_in])
            out[f"global_net/{l.name}/bias"] = (l.b, l.db)
        for l in [l for net in self.qt for l in net]:
            out[f"global_net/{l.name}/weight"] = (l.W[:l.k_in], None)
            out[f"global_net/{l.name}/bias"] = (l.b, None)
        out["global_net/actor/samples"] = (self.loc, self.dloc)
        out["global_net/actor/samples_std"] = (self.logstd, self.dlogstd)
        out["global_net/alpha/log_alpha"] = (self.log_alpha, self.dlog_alpha)
        return out

    def state_dict(self):
        sd = super().state_dict()
        sd["target_params"] = self.target_params.clone()
        sd["train_rng"] = int(self._train_rng.item())
        return sd

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        self.target_params.copy_(sd["target_params"])
        self._train_rng.fill_(int(sd.get("train_rng", 0)))
        self._gstep_dev.fill_(int(self.global_step))


class SACOptimizer:
    """clip_grads + Opt(c_optimizer, a_optimizer) of the SAC worker (base_worker.py:32-43,64-120): one joint
    clip_by_global_norm over the critic and actor gradient lists, AdamOptimizer(lr_critic) on the critic variables, then
    AdamOptimizer(lr_actor) on the actor variables and log_alpha, then the network's train_ops.  With more than one rank the
    clipped gradients and the pushed normaliser statistics are averaged first (SyncReplicasOptimizer semantics)."""

    def __init__(self, lr_critic=1e-4, lr_actor=1e-4, norm_clip=1.0, beta1=0.9, beta2=0.999, eps=1e-8, group=None):
        self.lr_critic, self.lr_actor, self.norm_clip = float(lr_critic), float(lr_actor), float(norm_clip or 0.0)
        self.beta1, self.beta2, self.eps, self.group = beta1, beta2, eps, group
        self.step = 0
        self.m: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None

    def apply_gradients(self, net: ParticleFilteringSACNetwork):
        self._launch_step(net)
        self._after_step(net)

    # Split as SyncReplicasAdam's step: `_launch_step` only enqueues work whose launch arguments never change (the Adam
    # step number lives in device memory), so GraphedSACUpdate can capture it; `_after_step` is host bookkeeping.
    def _launch_step(self, net: ParticleFilteringSACNetwork):
        import torch.distributed as dist
        st = _stream_ptr()
        if self.m is None:
            self.m, self.v = torch.zeros_like(net.params), torch.zeros_like(net.params)
            self.norm_scale = torch.zeros(2, dtype=torch.float32, device=net.params.device)
            self._scratch = torch.empty(296, dtype=torch.float64, device=net.params.device)
            self._step_dev = torch.full((1,), self.step, dtype=torch.int32, device=net.params.device)
            assert_replicas_identical(net.params, self.group)
        _cabi.check(_cabi.pfpn_clip_by_global_norm(net.grads.data_ptr(), net.n_params, self.norm_clip, self.norm_scale.data_ptr(),
                                                   self._scratch.data_ptr(), self._scratch.numel() * 8, st))
        # statistics pushed through the same accumulators as the gradients (normaliser moments of this minibatch and the
        # activity statistics: sync_model.py:37-45), then the mean over workers
        inv_n = 1.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            SyncReplicasAdam.pack_stats(self, net)
            dist.all_reduce(net.bucket, group=self.group)
            inv_n = 1.0 / dist.get_world_size(self.group)
            SyncReplicasAdam.unpack_stats(self, net, inv_n)
        elif net.normalize_state:  # one worker: the "mean over workers" of the pushed statistics is the statistics
            net.state_mean.copy_(net._new_mean)
            net.state_std.copy_(net._new_std)
        nc, n = net.n_critic, net.n_params
        for lo, hi, lr in ((0, nc, self.lr_critic), (nc, n, self.lr_actor)):  # step = device counter + 1
            _cabi.check(_cabi.pfpn_adam_step_dev(net.params[lo:hi].data_ptr(), net.grads[lo:hi].data_ptr(), self.m[lo:hi].data_ptr(),
                                                 self.v[lo:hi].data_ptr(), hi - lo, lr, self.beta1, self.beta2, self.eps,
                                                 self._step_dev.data_ptr(), 1, inv_n, st))
        self._step_dev.add_(1)
        net._gstep_dev.add_(1)
        if net.trainable:
            net.sync_target_net()  # (train_ops, sac.py:67-73; commutes with the particle resampling below)

    def _after_step(self, net: ParticleFilteringSACNetwork):
        self.step += 1
        net.global_step += 1
        for op in net.train_ops:
            if getattr(op, "__func__", None) is ParticleFilteringSACNetwork.sync_target_net:
                continue  # already enqueued by _launch_step
            op()


class GraphedSACUpdate:
    """One SAC learner step -- `net.compute_gradients(minibatch)` + `optimizer.apply_gradients(net)` -- captured ONCE in a
    CUDA graph and replayed.  At the reference's batch size (256, deepmimic_sac_base.py:8) the step is ~150 launch-latency
    sized kernels and copies: 2.5-2.9 ms eager, host-bound.  Capturable because the Philox offset of the training draws, the
    Adam step number and the normaliser's step are device words the kernels read when they run (`_train_rng`, `_step_dev`,
    `_gstep_dev`) and the graph itself advances.  The minibatch is copied into fixed input buffers (e.g. straight from
    ``ReplayRing.sample``); the particle-resampling tick (host-side interval logic) runs eagerly after the replay.
    Single-process use (with several ranks the NCCL all-reduce of the step stays eager: use apply_gradients)."""

    KEYS = ("state", "action", "reward", "not_terminal", "state_")

    def __init__(self, net: "ParticleFilteringSACNetwork", optimizer: "SACOptimizer", batch: int, warmup: int = 2):
        dev = net.device
        self.net, self.opt, self.B = net, optimizer, int(batch)
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.inputs = dict(state=f(batch, net.S), action=f(batch, net.A), reward=f(batch), not_terminal=f(batch),
                           state_=f(batch, net.S))
        self.graph, self.losses, self._warm, self.replays = None, None, int(warmup), 0

    def _set(self, *vals):
        for k, v in zip(self.KEYS, vals):
            if v is not None and v is not self.inputs[k]:
                self.inputs[k].copy_(torch.as_tensor(v, dtype=torch.float32).reshape(self.inputs[k].shape), non_blocking=True)

    def run(self, state=None, action=None, reward=None, not_terminal=None, state_=None):
        """Arguments left None keep what is already in ``self.inputs``.  Returns device scalars
        (loss, None, policy_loss, value_loss), valid until the next call."""
        import os
        self._set(state, action, reward, not_terminal, state_)
        net, opt = self.net, self.opt
        args = tuple(self.inputs[k] for k in self.KEYS)
        if self._warm > 0 or os.environ.get("PFPN_GRAPH", "1") == "0":  # eager steps (also: allocate buffers and workspaces)
            self._warm -= 1
            self.losses = net.compute_gradients(*args)
            opt.apply_gradients(net)
            return self.losses
        if self.graph is None:
            torch.cuda.synchronize(net.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.losses = net.compute_gradients(*args)
                opt._launch_step(net)
        self.graph.replay()
        self.replays += 1
        opt._after_step(net)
        return self.losses


class ReplayRing:
    """The DDPG/SAC worker's experience ring (models/workers/ddpg.py:11-27: `Buffer`, fixed capacity, overwrite at the
    write pointer) kept resident in device memory, one row per transition: state, action, reward, not_terminal, state_
    (`exp_buffers`, ddpg.py:44-46).  `sample(n)` is the off-policy minibatch draw of
    models/distributed_model.py:372-385 (`np.random.choice(len, n)`: uniform WITH replacement) as one gather."""

    FIELDS = ("state", "action", "reward", "not_terminal", "state_")

    def __init__(self, capacity: int, state_dim: int, action_dim: int, device="cuda", seed: int = 0):
        self.capacity, self.S, self.A = int(capacity), int(state_dim), int(action_dim)
        self.width = 2 * self.S + self.A + 2
        self.data = torch.empty(self.capacity, self.width, dtype=torch.float32, device=device)
        self.pointer = 0
        self.size = 0
        self.gen = torch.Generator(device=self.data.device)
        self.gen.manual_seed(seed)

    def __len__(self):
        return self.size

    def clear(self):
        self.pointer = self.size = 0

    def append(self, state, action, reward, not_terminal, state_):
        """Append one transition or a batch of them (leading dimension), wrapping around at the capacity."""
        dev = self.data.device
        f = lambda t, w: torch.as_tensor(t, dtype=torch.float32, device=dev).reshape(-1, w)
        rows = torch.cat([f(state, self.S), f(action, self.A), f(reward, 1), f(not_terminal, 1), f(state_, self.S)], dim=1)
        n = rows.shape[0]
        if n > self.capacity:  # only the last `capacity` rows survive, as with repeated single appends
            rows = rows[n - self.capacity:]
            self.pointer = (self.pointer + n - self.capacity) % self.capacity
            n = self.capacity
        first = min(n, self.capacity - self.pointer)
        self.data[self.pointer:self.pointer + first].copy_(rows[:first])
        if n > first:
            self.data[:n - first].copy_(rows[first:])
        self.pointer = (self.pointer + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, n: int):
        """(state [n,S], action [n,A], reward [n], not_terminal [n], state_ [n,S]) drawn uniformly with replacement."""
        if self.size == 0:
            raise ValueError("empty replay ring")
        idx = torch.randint(0, self.size, (n,), device=self.data.device, generator=self.gen)
        rows = self.data.index_select(0, idx)
        S, A = self.S, self.A
        return (rows[:, :S].contiguous(), rows[:, S:S + A].contiguous(), rows[:, S + A].contiguous(),
                rows[:, S + A + 1].contiguous(), rows[:, S + A + 2:].contiguous())
