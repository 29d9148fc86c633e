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

#define PFPN_ABI_VERSION 1

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
/* Minibatch advantage statistics: stats = {mean, 1/(sqrt(popvar)+1e-8)}
 * (actor_critic.py:151-155).  One CTA, deterministic. */
int pfpn_adv_stats(const float* adv, int32_t B, float* stats, pfpn_stream_t stream);

/* Introspection for tests / bench: SM count, resident CTAs per SM, CTA size and
 * states per tile the K1 launch for (A,P,mode) uses.  out[4]. */
int pfpn_head_launch_info(int32_t A, int32_t P, uint32_t mode, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* PFPN_B200_H_ */
