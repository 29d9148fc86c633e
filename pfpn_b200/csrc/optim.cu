// K7 and the small element-wise pieces of the learner update.
//
//  * state normaliser + clip              actor_critic.py:223-244  ((s - mean) / std, clip +-5)
//  * moving-average statistics update     networks/utils.py:60-68
//  * PPO value loss and its gradient      ppo.py:31-42, actor_critic.py:128-136
//  * clip_by_global_norm                  workers/base_worker.py:97-102 (TF semantics, [graph])
//  * Adam                                 workers/base_worker.py:64-70  (TF AdamOptimizer)
// Every reduction is two-stage with a fixed order, so the update is bit-reproducible.
#include "common.cuh"

namespace pfpn {

// x[b, 0:S] = clip((s - mean) / std, -c, c); columns S..ldo-1 are zero padding (GEMM K % 4 == 0)
__global__ void state_normalize_kernel(const float* __restrict__ s, const float* __restrict__ mean,
                                       const float* __restrict__ std, float* __restrict__ out, int B, int S, int ldo,
                                       float clip, int normalize) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * ldo) return;
  const int b = (int)(i / ldo), k = (int)(i % ldo);
  float v = 0.f;
  if (k < S) {
    v = s[(size_t)b * S + k];
    if (normalize) v = (v - mean[k]) / std[k];
    if (clip > 0.f) v = fminf(fmaxf(v, -clip), clip);
  }
  out[i] = v;
}

// column mean / population variance of X[B,S] (tf.nn.moments: mean, then mean of squared differences)
__global__ void __launch_bounds__(256) col_moments_kernel(const float* __restrict__ X, int B, int S,
                                                          float* __restrict__ mean_out, float* __restrict__ var_out) {
  const int k = blockIdx.x;
  __shared__ double sh[256];
  double s = 0.0;
  for (int b = threadIdx.x; b < B; b += 256) s += X[(size_t)b * S + k];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const double mean = sh[0] / B;
  __syncthreads();
  double q = 0.0;
  for (int b = threadIdx.x; b < B; b += 256) {
    const double d = X[(size_t)b * S + k] - mean;
    q += d * d;
  }
  sh[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    mean_out[k] = (float)mean;
    var_out[k] = (float)(sh[0] / B);
  }
}
// utils.py:60-68: decay = min(0.9999, (1+step)/(10+step)); mean/std moving averages, std >= 1e-6
__global__ void normalizer_update_kernel(float* __restrict__ mean, float* __restrict__ std, const float* __restrict__ m,
                                         const float* __restrict__ v, int S, float step) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  const float decay = fminf(0.9999f, (1.f + step) / (10.f + step));
  mean[k] = decay * mean[k] + (1.f - decay) * m[k];
  std[k] = fmaxf(1e-6f, decay * std[k] + (1.f - decay) * sqrtf(v[k]));
}

// Graph-capturable form of the statistics update (the step comes from DEVICE memory, inputs and outputs are separate
// buffers): coalesced partial column sums (thread = column, CTA = a block of rows; fp64), then one fixed-order finalize
// that also applies the moving average.  var = E[x^2] - mean^2 in fp64 (tf.nn.moments computes mean((x - mean)^2)).
constexpr int kMomentBlocks = 148;
__global__ void __launch_bounds__(256) moments_partial_kernel(const float* __restrict__ X, int B, int S, int rows_per,
                                                              double* __restrict__ part) {
  const int r0 = blockIdx.x * rows_per, r1 = min(B, r0 + rows_per);
  for (int k = threadIdx.x; k < S; k += 256) {
    double s = 0.0, q = 0.0;
    int b = r0;
    for (; b + 3 < r1; b += 4) {  // four independent loads in flight
      const float x0 = X[(size_t)b * S + k], x1 = X[(size_t)(b + 1) * S + k], x2 = X[(size_t)(b + 2) * S + k],
                  x3 = X[(size_t)(b + 3) * S + k];
      s += (double)x0 + (double)x1 + (double)x2 + (double)x3;
      q += (double)x0 * x0 + (double)x1 * x1 + (double)x2 * x2 + (double)x3 * x3;
    }
    for (; b < r1; ++b) {
      const float x = X[(size_t)b * S + k];
      s += x;
      q += (double)x * x;
    }
    part[((size_t)blockIdx.x * 2 + 0) * S + k] = s;
    part[((size_t)blockIdx.x * 2 + 1) * S + k] = q;
  }
}
__global__ void normalizer_finalize_kernel(const double* __restrict__ part, int nblk, int B, int S,
                                           const float* __restrict__ mean_in, const float* __restrict__ std_in,
                                           float* __restrict__ mean_out, float* __restrict__ std_out,
                                           const int* __restrict__ global_step) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  double s = 0.0, q = 0.0;
  for (int i0 = 0; i0 < nblk; i0 += 8) {  // loads batched eight partial pairs deep, additions in the same fixed order
    double xs[8], xq[8];                   // (one dependent L2 round trip per partial made this 34 us at 296 partials)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = i0 + u < nblk;
      xs[u] = ok ? __ldcg(&part[((size_t)(i0 + u) * 2 + 0) * S + k]) : 0.0;
      xq[u] = ok ? __ldcg(&part[((size_t)(i0 + u) * 2 + 1) * S + k]) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (i0 + u < nblk) {
        s += xs[u];
        q += xq[u];
      }
    }
  }
  const double mean = s / B;
  double var = q / B - mean * mean;
  if (var < 0.0) var = 0.0;
  const float step = (float)*global_step;
  const float decay = fminf(0.9999f, (1.f + step) / (10.f + step));
  mean_out[k] = decay * mean_in[k] + (1.f - decay) * (float)mean;
  std_out[k] = fmaxf(1e-6f, decay * std_in[k] + (1.f - decay) * sqrtf((float)var));
}

// value loss: mean((v - sg(adv + v_old))^2); dv = coef * 2 (v - target) * scale; one CTA
__global__ void __launch_bounds__(1024) value_loss_kernel(const float* __restrict__ v, const float* __restrict__ adv,
                                                          const float* __restrict__ v_old, float* __restrict__ dv,
                                                          float* __restrict__ loss, int B, float coef, float scale) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < B; i += 1024) {
    const float d = v[i] - (adv[i] + v_old[i]);
    s += (double)d * d;
    dv[i] = coef * 2.f * d * scale;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += sh[w];
    *loss = (float)(t * scale);  // un-weighted value loss (the reference reports it before * coef)
  }
}

// ---- SAC: every per-state scalar of the two losses in one launch ----------------------------------
// Reference: AbstractSACNetwork.build_q / setup_value_target_tensor / build_value_loss / build_policy_loss
// (networks/actor_critic/sac.py:107-126,132-139,160-173):
//   vf' = min(q1', q2')(s', a') - alpha logp(a'|s');  q_target = sg(r + gamma nt vf')
//   value_loss  = coef mean((q_target - q1(s,a_hist))^2 + (q_target - q2(s,a_hist))^2)
//   policy_loss = mean(alpha logp - min(q1, q2)(s, a) - log_alpha sg(logp + target_entropy)),  alpha = sg(exp(log_alpha))
// Outputs the gradients entering the four critic evaluations, the head (dL/dlogp) and log_alpha.
// tf.minimum routes the gradient to x where x <= y.  One CTA, fixed summation order.
__global__ void __launch_bounds__(1024) sac_losses_kernel(const float* __restrict__ q1a, const float* __restrict__ q2a,
                                                          const float* __restrict__ q1r, const float* __restrict__ q2r,
                                                          const float* __restrict__ q1t, const float* __restrict__ q2t,
                                                          const float* __restrict__ logp, const float* __restrict__ logp_t,
                                                          const float* __restrict__ reward, const float* __restrict__ not_terminal,
                                                          const float* __restrict__ log_alpha_p, float gamma, float coef,
                                                          float target_entropy, int B, float* __restrict__ dq1a,
                                                          float* __restrict__ dq2a, float* __restrict__ dq1r,
                                                          float* __restrict__ dq2r, float* __restrict__ dlogp,
                                                          float* __restrict__ out) {
  __shared__ double sh[3][32];
  const float log_alpha = *log_alpha_p;
  const float alpha = expf(log_alpha);
  const float inv_b = 1.f / (float)B;
  double sv = 0.0, sp = 0.0, sa = 0.0;
  for (int i = threadIdx.x; i < B; i += 1024) {
    const float vf = fminf(q1t[i], q2t[i]) - alpha * logp_t[i];
    const float qt = reward[i] + gamma * not_terminal[i] * vf;
    const float e1 = qt - q1r[i], e2 = qt - q2r[i];
    sv += (double)e1 * e1 + (double)e2 * e2;
    dq1r[i] = -2.f * coef * e1 * inv_b;
    dq2r[i] = -2.f * coef * e2 * inv_b;
    const float lp = logp[i];
    const bool first = q1a[i] <= q2a[i];
    sp += (double)(alpha * lp - (first ? q1a[i] : q2a[i]) - log_alpha * (lp + target_entropy));
    sa += (double)(lp + target_entropy);
    dq1a[i] = first ? -inv_b : 0.f;
    dq2a[i] = first ? 0.f : -inv_b;
    dlogp[i] = alpha * inv_b;
  }
  for (int o = 16; o > 0; o >>= 1) {
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = sv;
    sh[1][threadIdx.x >> 5] = sp;
    sh[2][threadIdx.x >> 5] = sa;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tv = 0.0, tp = 0.0, ta = 0.0;
    for (int w = 0; w < 32; ++w) {
      tv += sh[0][w];
      tp += sh[1][w];
      ta += sh[2][w];
    }
    out[0] = (float)(coef * tv * inv_b);   // value_loss (already times value_loss_coef, as self.value_loss is after `*=`)
    out[1] = (float)(tp * inv_b);          // policy_loss
    out[2] = (float)(-ta * inv_b);         // d policy_loss / d log_alpha
    out[3] = alpha;
  }
}
// y = a y + b x  (soft target update: sac.py:67-73 with a = 1 - tau, b = tau)
__global__ void axpby_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n, float a, float b) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = a * y[i] + b * x[i];
}

constexpr int kNormBlocks = 296;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, size_t n, double* __restrict__ part) {
  __shared__ double sh[8];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const double x = g[i];
    s += x * x;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    part[blockIdx.x] = t;
  }
}
// out[0] = global norm, out[1] = scale = clip * min(1/norm, 1/clip), NaN when the norm is not finite
__global__ void norm_finalize_kernel(const double* __restrict__ part, int nparts, float clip, float* __restrict__ out) {
  double t = 0.0;
  for (int i = 0; i < nparts; ++i) t += part[i];
  const float norm = (float)sqrt(t);
  out[0] = norm;
  float scale = 1.f;
  if (clip > 0.f) scale = isfinite(norm) ? clip * fminf(1.f / norm, 1.f / clip) : __int_as_float(0x7fc00000);
  out[1] = scale;
}
__global__ void scale_kernel(float* __restrict__ g, size_t n, const float* __restrict__ norm_scale) {
  const float s = norm_scale[1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) g[i] *= s;
}
// TF AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t); m,v moving averages; p -= lr_t m / (sqrt(v) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr_t, float b1, float b2, float eps, float gscale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// the same with the step number read from device memory (graph-capturable): lr_t as pfpn_adam_step computes it on the host
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                size_t n, float lr, float b1, float b2, float eps, float gscale,
                                const int* __restrict__ step_counter, int step_bias) {
  __shared__ float s_lr;
  if (threadIdx.x == 0) {
    const double t = (double)(*step_counter + step_bias);
    s_lr = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  __syncthreads();
  const float lr_t = s_lr;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// out[c, r] = in[r, c]  (weights -> K-major operand of the tensor-core forward GEMM)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc, int ldi, int ldo) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < R && c < Cc) ? in[(size_t)r * ldi + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) out[(size_t)c * ldo + r] = t[threadIdx.x][i];
  }
}

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_transpose(const float* in, int32_t ldi, float* out, int32_t ldo, int32_t rows, int32_t cols,
                              pfpn_stream_t stream_) {
  if (!in || !out || rows <= 0 || cols <= 0 || ldi < cols || ldo < rows) return PFPN_ERR_ARG;
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(stream_)>>>(in, out, rows, cols, ldi, ldo);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_state_normalize(const float* state, const float* mean, const float* std, float* out, int32_t B,
                                    int32_t S, int32_t ldo, float clip, int32_t normalize, pfpn_stream_t stream_) {
  if (!state || !out || B < 0 || S <= 0 || ldo < S) return PFPN_ERR_ARG;
  if (normalize && (!mean || !std)) return PFPN_ERR_ARG;
  if (B == 0) return PFPN_OK;
  const size_t n = (size_t)B * ldo;
  state_normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      state, mean, std, out, B, S, ldo, clip, normalize);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

// scratch: 2*S floats
extern "C" int pfpn_normalizer_update(const float* state, float* mean, float* std, int32_t B, int32_t S, float step,
                                      float* scratch, pfpn_stream_t stream_) {
  if (!state || !mean || !std || !scratch || B <= 0 || S <= 0) return PFPN_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  col_moments_kernel<<<S, 256, 0, st>>>(state, B, S, scratch, scratch + S);
  PFPN_CUDA_OK(cudaGetLastError());
  normalizer_update_kernel<<<(S + 255) / 256, 256, 0, st>>>(mean, std, scratch, scratch + S, S, step);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_normalizer_scratch_bytes(int32_t S, size_t* bytes) {
  if (!bytes || S <= 0) return PFPN_ERR_ARG;
  *bytes = (size_t)kMomentBlocks * 2 * S * sizeof(double);
  return PFPN_OK;
}
extern "C" int pfpn_normalizer_update_dev(const float* state, const float* mean_in, const float* std_in, float* mean_out,
                                          float* std_out, int32_t B, int32_t S, const int32_t* global_step, void* scratch,
                                          size_t scratch_bytes, pfpn_stream_t stream_) {
  if (!state || !mean_in || !std_in || !mean_out || !std_out || !global_step || !scratch || B <= 0 || S <= 0) return PFPN_ERR_ARG;
  if (scratch_bytes < (size_t)kMomentBlocks * 2 * S * sizeof(double)) return PFPN_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(scratch) & 7u) return PFPN_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int rows_per = (B + kMomentBlocks - 1) / kMomentBlocks;
  const int nblk = (B + rows_per - 1) / rows_per;
  double* part = reinterpret_cast<double*>(scratch);
  moments_partial_kernel<<<nblk, 256, 0, st>>>(state, B, S, rows_per, part);
  PFPN_CUDA_OK(cudaGetLastError());
  normalizer_finalize_kernel<<<(S + 127) / 128, 128, 0, st>>>(part, nblk, B, S, mean_in, std_in, mean_out, std_out, global_step);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_value_loss(const float* v, const float* adv, const float* v_old, float* dv, float* loss, int32_t B,
                               float coef, float scale, pfpn_stream_t stream_) {
  if (!v || !adv || !v_old || !dv || !loss || B <= 0) return PFPN_ERR_ARG;
  value_loss_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(v, adv, v_old, dv, loss, B, coef, scale);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

// ---- rollout side: generalised advantage estimate + value target for E trajectories of T steps -------
// Reference: A2CNetwork.generalized_advantage_estimate / value_target_estimate (networks/actor_critic/
// a2c.py:30-49) over `discount` (networks/utils.py:5-15): td_t = r_t + gamma v_{t+1} - v_t (fp32, numpy op
// order), adv_t = td_t + gae_gamma adv_{t+1} scanned backwards from 0; gae_gamma == 0 -> adv = td.
// One thread per trajectory: the scan is sequential in the reference too, so the result is bit-exact
// (explicit _rn operations: no FMA contraction).  Layout [E, T] / [E, T+1] row-major.
namespace pfpn {
__global__ void gae_kernel(const float* __restrict__ reward, const float* __restrict__ value, float* __restrict__ adv,
                           float* __restrict__ vtarget, int E, int T, float gamma, float gae_gamma) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float* r = reward + (size_t)e * T;
  const float* v = value + (size_t)e * (T + 1);
  float run = 0.f;
  float v_next = v[T];
  for (int t = T - 1; t >= 0; --t) {
    const float vt = v[t];
    const float td = __fsub_rn(__fadd_rn(r[t], __fmul_rn(gamma, v_next)), vt);
    run = gae_gamma != 0.f ? __fadd_rn(td, __fmul_rn(gae_gamma, run)) : td;
    adv[(size_t)e * T + t] = run;
    if (vtarget != nullptr) vtarget[(size_t)e * T + t] = __fadd_rn(vt, run);
    v_next = vt;
  }
}
}  // namespace pfpn

extern "C" int pfpn_gae(const float* reward, const float* value, float* adv, float* vtarget, int32_t E, int32_t T, float gamma,
                        float gae_gamma, pfpn_stream_t stream_) {
  if (!reward || !value || !adv || E < 0 || T < 0) return PFPN_ERR_ARG;
  if (E == 0 || T == 0) return PFPN_OK;
  pfpn::gae_kernel<<<(E + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(reward, value, adv, vtarget, E, T, gamma,
                                                                                         gae_gamma);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_sac_losses(const float* q1a, const float* q2a, const float* q1r, const float* q2r, const float* q1t,
                               const float* q2t, const float* logp, const float* logp_t, const float* reward,
                               const float* not_terminal, const float* log_alpha, float gamma, float coef, float target_entropy,
                               int32_t B, float* dq1a, float* dq2a, float* dq1r, float* dq2r, float* dlogp, float* out4,
                               pfpn_stream_t stream_) {
  if (!q1a || !q2a || !q1r || !q2r || !q1t || !q2t || !logp || !logp_t || !reward || !not_terminal || !log_alpha || !dq1a ||
      !dq2a || !dq1r || !dq2r || !dlogp || !out4 || B <= 0)
    return PFPN_ERR_ARG;
  pfpn::sac_losses_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      q1a, q2a, q1r, q2r, q1t, q2t, logp, logp_t, reward, not_terminal, log_alpha, gamma, coef, target_entropy, B, dq1a, dq2a,
      dq1r, dq2r, dlogp, out4);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_axpby(float* y, const float* x, size_t n, float a, float b, pfpn_stream_t stream_) {
  if (!y || !x) return PFPN_ERR_ARG;
  if (n == 0) return PFPN_OK;
  size_t grid = (n + 255) / 256;
  if (grid > 1184) grid = 1184;
  pfpn::axpby_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(y, x, n, a, b);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

// norm_scale[2] = {global norm, applied scale}; scratch: kNormBlocks doubles.  In place on `grads`.
extern "C" int pfpn_clip_by_global_norm(float* grads, size_t n, float clip, float* norm_scale, void* scratch,
                                        size_t scratch_bytes, pfpn_stream_t stream_) {
  if (!grads || !norm_scale || !scratch || n == 0) return PFPN_ERR_ARG;
  if (scratch_bytes < kNormBlocks * sizeof(double)) return PFPN_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(scratch) & 7u) return PFPN_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  double* part = reinterpret_cast<double*>(scratch);
  sumsq_partial_kernel<<<kNormBlocks, 256, 0, st>>>(grads, n, part);
  PFPN_CUDA_OK(cudaGetLastError());
  norm_finalize_kernel<<<1, 1, 0, st>>>(part, kNormBlocks, clip, norm_scale);
  PFPN_CUDA_OK(cudaGetLastError());
  if (clip > 0.f) {
    scale_kernel<<<kNormBlocks, 256, 0, st>>>(grads, n, norm_scale);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  return PFPN_OK;
}

extern "C" int pfpn_adam_step_dev(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                                  float beta2, float eps, const int32_t* step_counter, int32_t step_bias, float grad_scale,
                                  pfpn_stream_t stream_) {
  if (!params || !grads || !m || !v || !step_counter || n == 0) return PFPN_ERR_ARG;
  adam_dev_kernel<<<kNormBlocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(params, grads, m, v, n, lr, beta1, beta2, eps,
                                                                                     grad_scale, step_counter, step_bias);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_adam_step(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                              float beta2, float eps, int64_t step, float grad_scale, pfpn_stream_t stream_) {
  if (!params || !grads || !m || !v || n == 0 || step < 1) return PFPN_ERR_ARG;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  adam_kernel<<<kNormBlocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(params, grads, m, v, n, (float)lr_t, beta1,
                                                                                 beta2, eps, grad_scale);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}
