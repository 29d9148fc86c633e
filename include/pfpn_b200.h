/*
 * pfpn_b200 -- C ABI of the B200-native PFPN policy-head hot path.
 *
 * The reference (xupei0610/PFPN) has no FFI: its seam is the Python object
 * protocol of `MixtureGaussianDistribution` (networks/utils.py:85-236) and
 * `ParticleFilteringA2CNetwork` (networks/actor_critic/a2c.py:310-559).  Each
 * entry point below replaces the TF-1.14 sub-graph that the cited reference
 * lines build; `pfpn_b200/_cabi.py` is the ctypes binding, INTEGRATION.md shows
 * the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (fp32 unless noted,
 *    row-major, contiguous, 16-byte aligned); kernels never allocate;
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous;
 *  - return value: 0 = ok, <0 = argument error (PFPN_ERR_*), >0 = cudaError_t;
 *  - re-entrant, no global mutable state; one CUDA device per process.
 */
#ifndef PFPN_B200_H_
#define PFPN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: round 2 -- pfpn_rsample_args / pfpn_sac_head_args gained `offset_dev`, pfpn_head_push the packet protocol,
 *    pfpn_sync_args `params_lo`; new entry points are additive. */
#define PFPN_ABI_VERSION 2

typedef void* pfpn_stream_t;

enum pfpn_status {
  PFPN_OK = 0,
  PFPN_ERR_ARG = -1,         /* null pointer / non-positive size / bad enum   */
  PFPN_ERR_ALIGN = -2,       /* pointer not 16-byte aligned                   */
  PFPN_ERR_UNSUPPORTED = -3, /* shape outside the compiled instantiations     */
  PFPN_ERR_WORKSPACE = -4    /* workspace too small                           */
};

int pfpn_abi_version(void);
/* Static string for a status returned by any pfpn_* call. */
const char* pfpn_status_string(int status);

/* ------------------------------------------------------------------------
 * K1  mixture log_prob + categorical entropy (+ PPO surrogate) forward and
 *     backward in ONE pass over logits[B,A,P].
 * Replaces: MixtureGaussianDistribution.log_prob   networks/utils.py:108-144
 *           MixtureGaussianDistribution.entropy    networks/utils.py:146-151
 *           ClipPPONetwork.build_policy_loss       networks/actor_critic/ppo.py:44-54
 *           advantage normalisation                networks/actor_critic/actor_critic.py:151-155
 *           and their tf.gradients w.r.t. logits / samples / samples_std.
 * ---------------------------------------------------------------------- */
enum pfpn_head_mode {
  PFPN_HEAD_FWD = 0,  /* lp, ent only                                              */
  PFPN_HEAD_GRAD = 1, /* + vjp with caller-supplied dL/dlp[B] (and entropy grads)  */
  PFPN_HEAD_PPO = 2   /* + clipped-surrogate loss; dL/dlp derived in-kernel        */
};

#define PFPN_HEAD_FLAG_TANH 1u /* normalize_output=True: `value` is the pre-tanh u (utils.py:120-133) */

typedef struct pfpn_head_args {
  /* inputs */
  const float* logits;   /* [B, A, P]                                             */
  const float* loc;      /* [A, P]   particle means  ("samples")                  */
  const float* logstd;   /* [A, P]   particle log-std ("samples_std")             */
  const float* value;    /* [B, A]   action (pre-tanh when FLAG_TANH)              */
  const float* g_lp;     /* [B]      GRAD: dL/dlog_prob                            */
  const float* g_ent_ba; /* [B, A]   GRAD/PPO: dL/dentropy[b,a], may be NULL       */
  float g_ent;           /* scalar dL/dentropy[b,a] added to g_ent_ba (-beta/B)    */
  const float* adv;      /* [B]      PPO: raw advantage                            */
  const float* lp_old;   /* [B]      PPO: behaviour log_prob ("pi_running")        */
  const float* adv_stats;/* [2]      PPO: {mean, 1/(std+1e-8)} or NULL = identity  */
  float eps_clip;        /* PPO: epsilon (reference 0.2, ppo.py:16)                */
  float loss_scale;      /* PPO: 1/B_total -- the mean() of ppo.py:54              */
  /* outputs */
  float* lp;             /* [B]      sum_a log p                                   */
  float* ent;            /* [B]      sum_a H[b,a], may be NULL                     */
  float* ent_ba;         /* [B, A]   H[b,a], may be NULL                           */
  float* dlogits;        /* [B, A, P] GRAD/PPO (may alias logits)                  */
  float* dloc;           /* [A, P]   GRAD/PPO, overwritten                         */
  float* dlogstd;        /* [A, P]   GRAD/PPO, overwritten                         */
  float* dvalue;         /* [B, A]   GRAD only, may be NULL                        */
  float* loss;           /* [1]      PPO: -mean(min(surr, clipped)), overwritten   */
  int32_t B, A, P;
  uint32_t mode;         /* enum pfpn_head_mode                                    */
  uint32_t flags;        /* PFPN_HEAD_FLAG_*                                       */
} pfpn_head_args;

/* Bytes of scratch `pfpn_head_logprob` needs for (A, P) on the current device. */
int pfpn_head_workspace_bytes(int32_t A, int32_t P, size_t* bytes);
int pfpn_head_logprob(const pfpn_head_args* args, void* workspace, size_t workspace_bytes,
                      pfpn_stream_t stream);
/* Data-parallel form: the [2, A, P] particle gradients are the sharded head's only exchange.  With a `push`
 * the kernel that finishes dloc / dlogstd also writes them into row `rank` of EVERY rank's peer-mapped gather buffer
 * (NVLink stores) and its last CTA publishes `value` (the call counter: 1, 2, 3, ...) in every rank's flag word with a
 * system-scope release -- no separate exchange kernel sits behind K1.  The consumer calls pfpn_peer_gather_sum, which
 * waits for the N flags and sums the N LOCAL rows in rank order (identical on every rank).  args->dloc / dlogstd may be
 * NULL here.  Replaces (semantics) the accumulator sum of models/sync_model.py:92-96 for `samples` / `samples_std`. */
typedef struct pfpn_head_push {
  float* out[8];     /* out[p]: rank p's gather row for THIS rank and this call's slot (peer-mapped): 2*A*P floats         */
                     /*         (protocol 0) or 2*A*P 8-byte packets (protocol 1)                                        */
  int32_t* flags[8]; /* protocol 0: flags[p] = rank p's flag word for THIS rank (peer-mapped, zero-initialised, monotonic)*/
  int32_t* ticket;   /* protocol 0: local device int32, zero-initialised (CTA-arrival counter, self-resetting)            */
  int32_t nranks;    /* 1..8                                                                                              */
  int32_t value;     /* exchange counter, +1 per call (>= 1)                                                              */
  /* protocol 1 = PACKETS: every element travels as one 8-byte store {float value, int32 sequence = `value`}; an aligned
   * 8-byte store is single-copy atomic, so the receiver validates each element by its own sequence word -- no fence, no
   * ticket, no flag round trip behind the data (one NVLink write latency instead of three).  The same launch can also
   * CONSUME an earlier exchange: the thread that owns element i sums packet i of the N local rows of exchange
   * `consume_value` in rank order (spinning on the sequence words; with consume_value = value - 1 they arrived a step
   * ago) and writes consume_out[i] -- the exchange then costs no kernel of its own.  Gather memory must start zeroed.   */
  int32_t protocol;
  int32_t consume_value;     /* protocol 1: 0 = consume nothing in this launch                                            */
  const void* consume_rows;  /* protocol 1: this rank's LOCAL rows of exchange consume_value, [nranks][2*A*P] packets     */
  float* consume_out;        /* protocol 1: [2*A*P] = consume_scale * sum over ranks                                      */
  float consume_scale;
  int32_t reserved;
} pfpn_head_push;
int pfpn_head_logprob_push(const pfpn_head_args* args, void* workspace, size_t workspace_bytes,
                           const pfpn_head_push* push, pfpn_stream_t stream);
/* Stand-alone consumer of a protocol-1 exchange (the last one of a run, or a caller that needs the sum at once):
 * out[i] = scale * sum over ranks (rank order) of packet i of rows[r], waiting until every sequence word == value. */
int pfpn_peer_gather_sum_packets(const void* rows, int32_t nranks, int32_t value, size_t n, float* out, float scale,
                                 pfpn_stream_t stream);
/* out[n] = scale * sum_{r < nranks} gather[r*n + i] once flags[r] >= value for every r (acquire loads, system scope);
 * `gather` / `flags` are THIS rank's own buffers (the rows its peers pushed into).  n % 4 == 0. */
int pfpn_peer_gather_sum(const float* gather, const int32_t* flags, int32_t nranks, int32_t value, size_t n, float* out,
                         float scale, pfpn_stream_t stream);
/* Minibatch advantage statistics: stats = {mean, 1/(sqrt(popvar)+1e-8)}
 * (actor_critic.py:151-155).  One CTA, deterministic. */
int pfpn_adv_stats(const float* adv, int32_t B, float* stats, pfpn_stream_t stream);

/* Introspection for tests / bench: SM count, resident CTAs per SM, CTA size and
 * states per tile the K1 launch for (A,P,mode) uses.  out[4]. */
int pfpn_head_launch_info(int32_t A, int32_t P, uint32_t mode, int32_t* out);

/* ------------------------------------------------------------------------
 * K2  plain particle sampling (rollouts).
 * Replaces: MixtureGaussianDistribution.sample, plain branch   networks/utils.py:187-194
 *           (Categorical.sample -> TF Multinomial; Normal.sample; one_hot gather).
 * idx follows the TF-1.14 CPU Multinomial functor (fp64 CDF + upper_bound): with
 * ext_uniform supplied the particle indices are bit-exact against the oracle.
 * ---------------------------------------------------------------------- */
typedef struct pfpn_sample_args {
  const float* logits;      /* [B, A, P]                                            */
  const float* loc;         /* [A, P]                                               */
  const float* logstd;      /* [A, P]                                               */
  const double* ext_uniform;/* [B, A] fp64 in [0,1), or NULL = Philox                */
  const float* ext_normal;  /* [B, A, P] standard normals (read at [b,a,idx]) or NULL */
  float* action;            /* [B, A]   out                                          */
  int32_t* idx;             /* [B, A]   out: chosen particle ("dis_action")          */
  uint64_t seed, offset;    /* Philox key / call counter (production mode)           */
  int32_t B, A, P;
} pfpn_sample_args;
int pfpn_head_sample(const pfpn_sample_args* args, pfpn_stream_t stream);

/* K2f  the rollout side of the head in ONE pass over logits: particle index (TF Multinomial CPU semantics, bit-exact
 * with ext_uniform), action, mixture log_prob of that action, categorical entropy, and the running activity statistics
 * (max_active = max(max_active, max_b softmax), sum_active += sum_b softmax; both NULL to skip them).
 * Replaces: what ClipPPONetwork.run executes per step -- policy.sample + policy.log_prob + running_update_ops:
 *           networks/utils.py:187-194,108-151, networks/actor_critic/a2c.py:346-365, ppo.py:56-62.
 * Same Philox streams as pfpn_head_sample (uniform: (offset, row), normal: (offset + 1, row)).
 * Compiled for the shipped DPPO-PFPN shape A = 36, P = 35 (PFPN_ERR_UNSUPPORTED otherwise: use K2 + K1 + K4). */
typedef struct pfpn_rollout_args {
  const float* logits;       /* [B, A, P]                                            */
  const float* loc;          /* [A, P]                                               */
  const float* logstd;       /* [A, P]                                               */
  const double* ext_uniform; /* [B, A] fp64 in [0,1), or NULL = Philox                */
  const float* ext_normal;   /* [B, A, P] (read at [b,a,idx]) or NULL                 */
  float* action;             /* [B, A]  out                                          */
  int32_t* idx;              /* [B, A]  out: chosen particle ("dis_action")          */
  float* lp;                 /* [B]     out: log_prob(action)                        */
  float* ent;                /* [B]     out: sum_a H[b,a], may be NULL               */
  float* max_active;         /* [A, P]  in/out, may be NULL (with sum_active)        */
  float* sum_active;         /* [A, P]  in/out                                       */
  uint64_t seed, offset;
  int32_t B, A, P;
} pfpn_rollout_args;
int pfpn_rollout_workspace_bytes(int32_t A, int32_t P, size_t* bytes);
int pfpn_head_rollout(const pfpn_rollout_args* args, void* workspace, size_t workspace_bytes, pfpn_stream_t stream);

/* ------------------------------------------------------------------------
 * K3  reparameterised sampling (SAC, normalize_output=True) forward / backward.
 * Replaces: MixtureGaussianDistribution.sample, rsample branch  networks/utils.py:156-186
 *           incl. the custom gradients mask2 (:164-171) and mask (:176-183), and TFP 0.7
 *           RelaxedOneHotCategorical(1.0, logits).sample (Gumbel-softmax).
 * fwd: sample = tanh(s_pre), s_pre = (loc + scale*eps)[argmax softmax(logits + Gumbel(U))].
 * bwd: given dL/dsample and dL/ds_pre writes dlogits (overwritten) and ADDS into
 *      dloc / dlogstd (caller zeroes them).  The same (seed, offset) or the same ext_* arrays
 *      must be passed to fwd and bwd: the draws are regenerated, not stored.
 * ---------------------------------------------------------------------- */
typedef struct pfpn_rsample_args {
  const float* logits;      /* [B, A, P]                                            */
  const float* loc;         /* [A, P]                                               */
  const float* logstd;      /* [A, P]                                               */
  const float* ext_uniform; /* [B, A, P] in [tiny, 1) or NULL = Philox               */
  const float* ext_normal;  /* [B, A, P] or NULL (both or neither)                   */
  float* sample;            /* [B, A]  fwd out: tanh(s_pre)                          */
  float* s_pre;             /* [B, A]  fwd out: pre-tanh value                       */
  int32_t* idx;             /* [B, A]  fwd out: argmax particle ("dis_action")       */
  const float* g_sample;    /* [B, A]  bwd in                                        */
  const float* g_s_pre;     /* [B, A]  bwd in, may be NULL (= 0)                     */
  float* dlogits;           /* [B, A, P] bwd out                                     */
  float* dloc;              /* [A, P]  bwd in/out (accumulated)                      */
  float* dlogstd;           /* [A, P]  bwd in/out (accumulated)                      */
  uint64_t seed, offset;
  int32_t B, A, P;
  int32_t reserved;
  const uint64_t* offset_dev; /* optional DEVICE word added to `offset` when the kernel starts (NULL = 0): a captured CUDA   */
                              /* graph then draws fresh variates on every replay (the caller advances the word in-graph)     */
} pfpn_rsample_args;
int pfpn_head_rsample_fwd(const pfpn_rsample_args* args, pfpn_stream_t stream);
/* (backward: dloc / dlogstd are ADDED to -- zero or pre-load them; the accumulation runs in per-lane registers and fixed-order
 *  partials, so it is bit-reproducible; workspace from pfpn_rsample_bwd_workspace_bytes) */
int pfpn_rsample_bwd_workspace_bytes(int32_t B, int32_t A, int32_t P, size_t* bytes);
int pfpn_head_rsample_bwd(const pfpn_rsample_args* args, void* workspace, size_t workspace_bytes, pfpn_stream_t stream);

/* K3f  the SAC-PFPN head in ONE pass (the fused fwd_bwd of the c5 sweep): reparameterised sample, the
 * tanh-squashed log_prob of that sample, and the backward of both given dL/dsample (from the critics) and dL/dlog_prob
 * -- 8AP + 12A + 8 bytes per state instead of the five [B,A,P] passes of the three-launch form, every draw generated
 * once.  With the same (seed, offset) it draws the same Gumbel / normal variates as pfpn_head_rsample_fwd / _bwd; with
 * ext_uniform / ext_normal (verification) the argmax particle is bit-exact against the oracle.
 * Replaces: MixtureGaussianDistribution.sample (rsample branch) + .log_prob((sample, s_pre)) + tf.gradients through both,
 *           networks/utils.py:108-144,156-186; AbstractSACNetwork.build_policy_loss's head-facing part, sac.py:166-173.
 * Compiled for the BASELINE c5 shape A = 36, P = 100 (PFPN_ERR_UNSUPPORTED otherwise: use the three-launch form). */
typedef struct pfpn_sac_head_args {
  const float* logits;      /* [B, A, P]                                              */
  const float* loc;         /* [A, P]                                                 */
  const float* logstd;      /* [A, P]                                                 */
  const float* ext_uniform; /* [B, A, P] in [tiny, 1) or NULL = Philox                 */
  const float* ext_normal;  /* [B, A, P] or NULL (both or neither)                     */
  const float* g_sample;    /* [B, A]  dL/dsample                                      */
  const float* g_lp;        /* [B]     dL/dlog_prob                                    */
  float* sample;            /* [B, A]  out: tanh(s_pre)                                */
  float* s_pre;             /* [B, A]  out                                             */
  int32_t* idx;             /* [B, A]  out: argmax particle                            */
  float* logp;              /* [B]     out: log_prob((sample, s_pre))                  */
  float* dlogits;           /* [B, A, P] out (may alias logits)                        */
  float* dloc;              /* [A, P]  out, overwritten                                */
  float* dlogstd;           /* [A, P]  out, overwritten                                */
  uint64_t seed, offset;
  int32_t B, A, P;
  int32_t reserved;
  const uint64_t* offset_dev; /* optional device word added to `offset` (see pfpn_rsample_args)                         */
} pfpn_sac_head_args;
int pfpn_sac_head_workspace_bytes(int32_t A, int32_t P, size_t* bytes);
int pfpn_sac_head_fwd_bwd(const pfpn_sac_head_args* args, void* workspace, size_t workspace_bytes, pfpn_stream_t stream);
/* Deterministic second stage over per-CTA [nparts][2*AP] partials in K1's convention: dloc = sum / exp(logstd), dlogstd = sum. */
int pfpn_head_finalize_partials(const float* part, int32_t nparts, const float* logstd, float* dloc, float* dlogstd,
                                int32_t AP, pfpn_stream_t stream);

/* Deterministic action (evaluator): MixtureGaussianDistribution.mean  networks/utils.py:202-236.
 * action[b,a] = loc[a, argmax_k logits[b,a,k]]  (tanh of it with PFPN_HEAD_FLAG_TANH). idx may be NULL. */
int pfpn_head_mean(const float* logits, const float* loc, float* action, int32_t* idx, int32_t B, int32_t A,
                   int32_t P, uint32_t flags, pfpn_stream_t stream);

/* ------------------------------------------------------------------------
 * K4  running activity statistics.
 * Replaces: ParticleFilteringA2CNetwork.init, `resample` scope   networks/actor_critic/a2c.py:346-365
 *   max_active = max(max_active, max_b softmax(logits)); sum_active += sum_b softmax(logits).
 * probs [B,A,P] may be NULL (only needed when the caller wants `dis_dist.probs`).
 * Deterministic (register accumulators per (a,k), ordered two-stage combine; no float atomics); P <= 256.
 * ---------------------------------------------------------------------- */
int pfpn_stats_workspace_bytes(int32_t B, int32_t A, int32_t P, size_t* bytes);
int pfpn_stats_update(const float* logits, float* probs, float* max_active, float* sum_active, int32_t B,
                      int32_t A, int32_t P, void* workspace, size_t workspace_bytes, pfpn_stream_t stream);

/* ------------------------------------------------------------------------
 * K5  dead-particle resampling.
 * Replaces: ParticleFilteringA2CNetwork.build_resample_ops      networks/actor_critic/a2c.py:385-474
 *           and the statistics reset of `update()`               a2c.py:370-378.
 * All tensors are updated in place.  The out_* pointers are optional verification outputs
 * (bit-exact integers against the oracle when the ext_* draws are supplied).
 * ---------------------------------------------------------------------- */
#define PFPN_RESAMPLE_FLAG_TANH 1u /* normalize_policy_output_: atanh(clip(loc)) (a2c.py:448-450) */
typedef struct pfpn_resample_args {
  float* max_active;         /* [A, P] in; zeroed on return                           */
  float* sum_active;         /* [A, P] in; zeroed on return                           */
  float* loc;                /* [A, P] in/out  ("samples")                            */
  float* logstd;             /* [A, P] in/out  ("samples_std")                        */
  float* bias;               /* [A*P]  in/out  (fc_policy/bias)                       */
  float* weight;             /* [H, A*P] in/out (fc_policy/weight), NULL iff H == 0   */
  const double* ext_cat_u;   /* [A, P] fp64 draws of tf.random.categorical, or NULL   */
  const int32_t* ext_choice; /* [A*P]  draws of uniform{0..k-1} (resample > 0), or NULL */
  const float* ext_noise_u;  /* [A*P]  draws of uniform(-1,1), or NULL                */
  int32_t* out_M;            /* [1]    number of dead particles                       */
  int32_t* out_nuniq;        /* [1]    number of distinct source columns              */
  int32_t* out_invalid;      /* [A*P, 2] (a_m, j_m) row-major                         */
  int32_t* out_cand;         /* [A, k] candidate table                                */
  int32_t* out_src;          /* [A*P]                                                 */
  int32_t* out_col;          /* [A*P]  a_m*P + j_m                                    */
  int32_t* out_tcol;         /* [A*P]  a_m*P + src_m                                  */
  int32_t* out_uniq;         /* [A*P]  unique_with_counts: values, first-occurrence order */
  int32_t* out_idx;          /* [A*P]  unique_with_counts: index of tcol_m in uniq    */
  int32_t* out_count;        /* [A*P]  unique_with_counts: counts                     */
  int32_t* out_delta;        /* [A*P]  1 iff uniq_u is itself a dead column           */
  uint64_t seed, offset;     /* Philox key / call counter (production mode)           */
  float threshold;           /* <= 0: 0.05 / P (a2c.py:391)                           */
  int32_t A, P, H;
  int32_t resample;          /* -1: categorical candidates (shipped); r > 0: top-r    */
  uint32_t flags;
} pfpn_resample_args;
int pfpn_resample_workspace_bytes(int32_t A, int32_t P, int32_t H, size_t* bytes);
int pfpn_resample(const pfpn_resample_args* args, void* workspace, size_t workspace_bytes,
                  pfpn_stream_t stream);

/* ------------------------------------------------------------------------
 * K6  dense layers of the 1024-512 trunk (fp32 FFMA anchor path).
 * Replaces: fc_layer                         networks/ops.py:82-118
 *           build_conv_fc_net                networks/utils.py:17-43
 *           and the MatMul_grad / Relu6Grad nodes tf.gradients derives from them.
 * Row-major fp32; every leading dimension and every contiguous extent (K, N) must be a
 * multiple of 4, except N == 1 (critic/fc3) which has dedicated paths.
 * ---------------------------------------------------------------------- */
/* Y[M,N] = act(X[M,K] W[K,N] + b[N]), act = relu6 if `relu6` else identity */
int pfpn_mlp_linear_fwd(const float* X, int32_t ldx, const float* W, const float* b, float* Y, int32_t ldy,
                        int32_t M, int32_t K, int32_t N, int32_t relu6, pfpn_stream_t stream);
/* dX[M,K] = (dY[M,N] W[K,N]^T) .* 1[0 < Hin < 6]   (Hin = forward output of the previous layer, NULL: no mask) */
int pfpn_mlp_linear_bwd_input(const float* dY, int32_t ldy, const float* W, const float* Hin, float* dX,
                              int32_t ldx, int32_t M, int32_t K, int32_t N, pfpn_stream_t stream);
/* dW[K,N] = X[M,K]^T dY[M,N], db[N] = column sums of dY; deterministic split-K over the batch */
int pfpn_mlp_wgrad_workspace_bytes(int32_t M, int32_t K, int32_t N, size_t* bytes);
int pfpn_mlp_linear_bwd_weight(const float* X, int32_t ldx, const float* dY, int32_t ldy, float* dW, float* db,
                               int32_t M, int32_t K, int32_t N, void* workspace, size_t workspace_bytes,
                               pfpn_stream_t stream);

/* K6, tensor-core path: 3xTF32 error-compensated GEMMs on tcgen05 (TMA-fed, TMEM accumulators), same
 * epilogues as the FFMA anchor.  Operands are read as stored -- K-major or MN-major UMMA descriptors --
 * so neither the weights nor the batch-major activations are ever transposed in HBM.
 * epi: 0 none, 1 +bias[n], 2 relu6(+bias), 3 multiply by 1[0 < Hm[m,n] < 6].
 *   pfpn_tc_gemm_nt: C[M,N] = epi(A[M,K] * Bt[N,K]^T)   input gradient dX = dY * W^T with Bt = W[in,out] as stored
 *                    (replaces the MatMul(transpose_b) of the fc_layer gradient, networks/ops.py:82-118)
 *   pfpn_tc_gemm_nn: C[M,N] = epi(A[M,K] * B[K,N])      forward act(X * W + b) with B = W as stored (ops.py:108-116) */
int pfpn_tc_gemm_nt(const float* A, int32_t lda, const float* Bt, int32_t ldb, float* C, int32_t ldc,
                    const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K,
                    int32_t epi, pfpn_stream_t stream);
int pfpn_tc_gemm_nn(const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                    const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K,
                    int32_t epi, pfpn_stream_t stream);
/* The same two GEMMs with the low part of the B operand precomputed: B_lo = B - tf32(B) (pfpn_split_lo), same layout
 * and ldb.  The weights are constant within an optimizer step, so they are split ONCE per step (the optimizer kernel of
 * pfpn_sync_step writes the low parts next to the parameters) instead of once per tile per GEMM by the splitter warps;
 * B_lo then arrives by TMA.  Bit-identical results. */
int pfpn_tc_gemm_nt_lo(const float* A, int32_t lda, const float* Bt, const float* Bt_lo, int32_t ldb, float* C, int32_t ldc,
                       const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K, int32_t epi,
                       pfpn_stream_t stream);
int pfpn_tc_gemm_nn_lo(const float* A, int32_t lda, const float* B, const float* B_lo, int32_t ldb, float* C, int32_t ldc,
                       const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K, int32_t epi,
                       pfpn_stream_t stream);
/* lo[i] = x[i] - tf32(x[i]), n % 4 == 0, 16-byte aligned. */
int pfpn_split_lo(const float* x, float* lo, size_t n, pfpn_stream_t stream);
/* out[cols, ldo] = in[rows, ldi]^T (utility; the GEMMs above no longer need transposed copies). */
int pfpn_transpose(const float* in, int32_t ldi, float* out, int32_t ldo, int32_t rows, int32_t cols,
                   pfpn_stream_t stream);
/* Weight gradient on tcgen05: dW[K,N] = X[M,K]^T dY[M,N], X and dY row-major as stored (both MN-major
 * operands; split-K over the batch in chunks of 2048 rows, deterministic fp32 second stage).  `db`
 * (nullable) receives the bias gradient db[N] = column sums of dY, accumulated from the dY tiles the
 * GEMM stages anyway (no extra pass over dY).  Replaces tf.gradients through fc_layer, networks/ops.py:108-116. */
int pfpn_tc_wgrad_workspace_bytes(int32_t M, int32_t K, int32_t N, size_t* bytes);
int pfpn_tc_linear_bwd_weight(const float* X, int32_t ldx, const float* dY, int32_t ldy, float* dW, float* db,
                              int32_t M, int32_t K, int32_t N, void* workspace, size_t workspace_bytes,
                              pfpn_stream_t stream);
/* db[N] = column sums of dY[M,N]; workspace >= 1024*N floats. */
int pfpn_bias_grad(const float* dY, int32_t ldy, float* db, int32_t M, int32_t N, void* workspace,
                   size_t workspace_bytes, pfpn_stream_t stream);

/* ------------------------------------------------------------------------
 * Learner-update element-wise pieces and K7 (clip + Adam).
 * ---------------------------------------------------------------------- */
/* out[B, ldo] = clip((state - mean) / std, +-clip), zero padded beyond S.
 * Replaces: build_state_normalizer_op        networks/actor_critic/actor_critic.py:223-244 */
int pfpn_state_normalize(const float* state, const float* mean, const float* std, float* out, int32_t B,
                         int32_t S, int32_t ldo, float clip, int32_t normalize, pfpn_stream_t stream);
/* Moving-average update of the state statistics over a minibatch; scratch = 2*S floats.
 * Replaces: online_normalizer(moving_average=True)   networks/utils.py:60-68 */
int pfpn_normalizer_update(const float* state, float* mean, float* std, int32_t B, int32_t S, float step,
                           float* scratch, pfpn_stream_t stream);
/* Graph-capturable form: the step (the network's global_step) is read from DEVICE memory, inputs and outputs are
 * separate buffers (the pushed statistics are computed from the pre-update values and applied after aggregation:
 * LocalUpdateHookPre, models/sync_model.py:123-138).  Coalesced fp64 partial sums, fixed order. */
int pfpn_normalizer_scratch_bytes(int32_t S, size_t* bytes);
int pfpn_normalizer_update_dev(const float* state, const float* mean_in, const float* std_in, float* mean_out,
                               float* std_out, int32_t B, int32_t S, const int32_t* global_step, void* scratch,
                               size_t scratch_bytes, pfpn_stream_t stream);
/* loss[0] = scale * sum((v - (adv + v_old))^2); dv = coef * 2 (v - target) * scale  (scale = 1/B).
 * Replaces: ClipPPONetwork.build_value_loss / setup_value_target_tensor   ppo.py:31-42 */
int pfpn_value_loss(const float* v, const float* adv, const float* v_old, float* dv, float* loss, int32_t B,
                    float coef, float scale, pfpn_stream_t stream);
/* Rollout side: generalised advantage estimate and value target of E trajectories of T steps,
 * reward [E,T], value [E,T+1] (bootstrap value last): td_t = r_t + gamma v_{t+1} - v_t, adv_t = td_t + gae_gamma adv_{t+1}
 * (gae_gamma = gamma*lambda; 0 -> adv = td), vtarget = v + adv (nullable).  Sequential fp32 scan per trajectory in the
 * reference's op order -> bit-exact.
 * Replaces: A2CNetwork.generalized_advantage_estimate / value_target_estimate  networks/actor_critic/a2c.py:30-49,
 *           discount  networks/utils.py:5-15 */
int pfpn_gae(const float* reward, const float* value, float* adv, float* vtarget, int32_t E, int32_t T, float gamma,
             float gae_gamma, pfpn_stream_t stream);
/* SAC learner step: every per-state scalar of the two losses in one launch.
 *   vf' = min(q1t, q2t) - alpha logp_t;  q_target = reward + gamma not_terminal vf'      (target net, stop-gradient)
 *   value_loss  = coef mean((q_target - q1r)^2 + (q_target - q2r)^2)                    (critics at the replayed action)
 *   policy_loss = mean(alpha logp - min(q1a, q2a) - log_alpha (logp + target_entropy))  (critics at the sampled action)
 * alpha = exp(*log_alpha) (device scalar, treated as a constant).  Writes d/dq1a, d/dq2a (tf.minimum: gradient to x where
 * x <= y), d/dq1r, d/dq2r, d policy_loss/d logp, and out4 = {value_loss, policy_loss, d policy_loss/d log_alpha, alpha}.
 * Replaces: AbstractSACNetwork.build_q / setup_value_target_tensor / build_value_loss / build_policy_loss
 *           networks/actor_critic/sac.py:107-126,132-139,160-173 */
int pfpn_sac_losses(const float* q1a, const float* q2a, const float* q1r, const float* q2r, const float* q1t,
                    const float* q2t, const float* logp, const float* logp_t, const float* reward,
                    const float* not_terminal, const float* log_alpha, float gamma, float coef, float target_entropy,
                    int32_t B, float* dq1a, float* dq2a, float* dq1r, float* dq2r, float* dlogp, float* out4,
                    pfpn_stream_t stream);
/* y = a*y + b*x: the soft target-network update v_ <- (1-tau) v_ + tau v   (sac.py:67-73) */
int pfpn_axpby(float* y, const float* x, size_t n, float a, float b, pfpn_stream_t stream);
/* grads *= clip * min(1/||grads||, 1/clip) (NaN if the norm is not finite); norm_scale = {norm, scale};
 * clip <= 0 only computes the norm.  scratch >= 296 doubles.
 * Replaces: clip_grads -> tf.clip_by_global_norm      models/workers/base_worker.py:97-102 */
int pfpn_clip_by_global_norm(float* grads, size_t n, float clip, float* norm_scale, void* scratch,
                             size_t scratch_bytes, pfpn_stream_t stream);
/* TF AdamOptimizer step `step` (1-based) on a flat parameter buffer; the gradient is read as
 * grads * grad_scale (1/N after a sum all-reduce over N ranks).
 * Replaces: build_optimizer -> tf.train.AdamOptimizer  models/workers/base_worker.py:64-70 */
int pfpn_adam_step(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                   float beta2, float eps, int64_t step, float grad_scale, pfpn_stream_t stream);
/* Same with the step number read from DEVICE memory when the kernel runs: step = *step_counter + step_bias (>= 1), lr_t
 * computed on the device -- nothing in the launch arguments changes between steps, so the call can be captured in a CUDA
 * graph (the caller advances the counter in-graph, e.g. after the last Adam launch of the step). */
int pfpn_adam_step_dev(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                       float beta2, float eps, const int32_t* step_counter, int32_t step_bias, float grad_scale,
                       pfpn_stream_t stream);

/* ------------------------------------------------------------------------
 * K7 fused with the data-parallel exchange: sum of the N ranks' clipped [gradient | statistics]
 * buckets read straight from peer memory (NVLink P2P) + mean + Adam in ONE kernel.
 * Replaces (semantics): SyncReplicasOptimizer accumulator mean + ApplyAdam
 *                       models/sync_model.py:92-96, models/workers/base_worker.py:64-70.
 * `buckets` / `flags` are HOST arrays of nranks DEVICE pointers (this rank's own buffer at index
 * `rank`, the others mapped through CUDA IPC); flags[r] is an int32[nranks] array owned by rank r.
 * Per step: write the bucket, pfpn_peer_signal(value = step), pfpn_peer_allreduce_adam(value = step).
 * ---------------------------------------------------------------------- */
int pfpn_enable_peer_access(int32_t peer_device);
/* Peer-visible buffers: cudaMalloc'ed + zeroed, exported as a 64-byte CUDA IPC handle; `pfpn_peer_open`
 * maps a peer's buffer into the caller's CURRENT device context with lazy peer access. */
int pfpn_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int pfpn_peer_open(const unsigned char* handle64, void** ptr);
int pfpn_peer_close(void* ptr);
int pfpn_peer_free(void* ptr);
int pfpn_peer_signal(const float* const* buckets, int32_t* const* flags, int32_t rank, int32_t nranks,
                     int32_t value, pfpn_stream_t stream);
int pfpn_peer_allreduce_adam(const float* const* buckets, int32_t* const* flags, int32_t rank, int32_t nranks,
                             int32_t value, size_t n_params, size_t n_total, float* params, float* m, float* v,
                             float* avg_out, float lr, float beta1, float beta2, float eps, int64_t step,
                             pfpn_stream_t stream);
/* Two-phase form for N >= 3 GPUs (reduce-scatter + all-gather inside ONE kernel; 2(N-1)/N instead of N-1 bucket
 * volumes per GPU over NVLink): rank r averages its 1/N slice in rank order, publishes it in `reduced[r]`, every rank
 * gathers the N averaged slices and applies Adam to all parameters.  `reduced`: HOST array of nranks peer-mapped
 * buffers of n_total floats; the flag buffers must hold >= 64 zero-initialised ints; `value` = 1, 2, 3, ... per call. */
int pfpn_peer_allreduce_adam_rs(const float* const* buckets, float* const* reduced, int32_t* const* flags, int32_t rank,
                                int32_t nranks, int32_t value, size_t n_params, size_t n_total, float* params, float* m,
                                float* v, float* avg_out, float lr, float beta1, float beta2, float eps, int64_t step,
                                pfpn_stream_t stream);
/* The sharded head's only exchange: out[n] = scale * sum_r buckets[r][n] for the [2,A,P] particle
 * gradients -- signal, wait and the rank-ordered sum in ONE kernel over peer memory (same staging / flag
 * protocol as above; `value` = call counter, +1 per call, buffers alternate by its parity). */
int pfpn_peer_allreduce_sum(const float* const* buckets, int32_t* const* flags, int32_t rank, int32_t nranks,
                            int32_t value, size_t n, float* out, float scale, pfpn_stream_t stream);

/* ------------------------------------------------------------------------
 * The whole data-parallel optimizer step, graph-capturable: local clip_by_global_norm -> this rank's scaled gradients and
 * the four pushed statistics staged in peer-visible memory -> flags -> rank-ordered sum over the N ranks out of NVLink
 * peer memory -> mean -> Adam -> averaged statistics assigned.  THREE launches; nothing in the arguments changes from
 * step to step: the exchange call number (flag value / staging parity), the Adam step and the network's global_step live
 * in DEVICE memory (`counters`, int32[4] = {calls, adam_step, global_step, ticket}; the first kernel increments the
 * first three), so one CUDA graph of the update can be replayed.
 * Replaces: clip_grads (models/workers/base_worker.py:97-102) + SyncReplicasOptimizer accumulators
 *           (models/sync_model.py:37-45,92-96) + ApplyAdam (base_worker.py:64-70).
 * Statistics tail of the bucket: [new_mean S][new_std S][max_active AP][sum_active AP]; S == 0 / AP == 0 drop a part.
 * nranks == 1: no peers, gradients scaled in place.  two_phase != 0: reduce-scatter + all-gather form (N >= 3).
 * ---------------------------------------------------------------------- */
typedef struct pfpn_sync_args {
  float* grads;            /* [n_total] this rank's bucket: gradients in [0, n_params)                           */
  size_t n_params, n_total;
  float clip;              /* <= 0: norm only                                                                    */
  const float* new_mean;   /* [S] this minibatch's moving-average state statistics (pushed)                      */
  const float* new_std;    /* [S]                                                                                */
  float* state_mean;       /* [S] receives the accumulator mean                                                  */
  float* state_std;        /* [S]                                                                                */
  float* max_active;       /* [AP] pushed AND assigned (averaged, not max-reduced: a reference quirk)            */
  float* sum_active;       /* [AP]                                                                               */
  int32_t S, AP;
  float* params;           /* [n_params]                                                                         */
  float* m;                /* [n_params] Adam slots                                                              */
  float* v;
  float lr, beta1, beta2, eps;
  int32_t* counters;       /* device int32[4], see above                                                         */
  float* norm_scale;       /* [2] out: {global norm, applied scale}                                              */
  void* scratch;           /* pfpn_sync_step_scratch_bytes, 8-byte aligned                                       */
  size_t scratch_bytes;
  float* const* stage;     /* HOST array [nranks]: rank r's staging base, 2 * n_total floats (peer-mapped)       */
  float* const* reduced;   /* HOST array [nranks]: rank r's averaged-slice buffer, n_total floats (two_phase)    */
  int32_t* const* flags;   /* HOST array [nranks]: rank r's flag words, int32[64], zero-initialised              */
  int32_t rank, nranks, two_phase;
  float* params_lo;        /* [n_params] or NULL: receives params - tf32(params) for the tensor-core GEMMs         */
} pfpn_sync_args;
int pfpn_sync_step_scratch_bytes(size_t* bytes);
int pfpn_sync_step(const pfpn_sync_args* args, pfpn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PFPN_B200_H_ */
