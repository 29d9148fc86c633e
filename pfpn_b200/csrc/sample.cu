// K2 -- plain particle sampling (DPPO rollouts), K3 -- reparameterised sampling forward /
// backward (SAC), and the deterministic action (`mean`).
//
// Reference semantics:
//   plain   /root/reference/networks/utils.py:187-194 : idx ~ Categorical(logits) via TF's
//           Multinomial op, action = (loc + scale * eps)[idx].  The TF-1.14 CPU functor builds an
//           fp64 running CDF of exp(double(logit) - max) and takes upper_bound(cdf, u * total);
//           that is reproduced literally (sequential fp64 accumulation) so that, given the same
//           uniforms, the particle indices are bit-exact.
//   rsample utils.py:156-186 : w = softmax(logits + Gumbel(U)) (TFP 0.7 RelaxedOneHotCategorical,
//           T = 1), k* = argmax w, s_ = (loc + scale*eps)[k*], sample = tanh(s_), with the two
//           straight-through custom gradients `mask2` / `mask`.
//   mean    utils.py:202-236.
// One warp owns one mixture row (b, a); lanes stride over the P particles.
// Randomness: verification mode reads caller-supplied uniforms / normals; production mode draws
// from Philox4x32-10 keyed by `seed`, counter (offset, row * P + k).
#include "common.cuh"

namespace pfpn {

constexpr int kSampleWarps = 8;
constexpr float kF32Tiny = 1.17549435e-38f;  // TFP: uniform(minval=finfo(f32).tiny, maxval=1)

__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Box-Muller on two 24-bit uniforms from one Philox block (returns one normal per call site)
__device__ __forceinline__ float philox_normal(const Philox& rng, uint64_t offset, uint64_t ctr) {
  const uint4 r = rng(offset, ctr);
  const float u1 = u32_to_unit_open(r.x), u2 = u32_to_unit_open(r.y);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// ------------------------------------------------------------------------------------------------
// K2: plain sample
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSampleWarps * 32) sample_kernel(const pfpn_sample_args ar) {
  extern __shared__ double cdf_s[];  // [kSampleWarps][P]
  const int A = ar.A, P = ar.P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* cdf = cdf_s + (size_t)warp * P;
  const long long rows = (long long)ar.B * A;
  const Philox rng(ar.seed);
  for (long long r = (long long)blockIdx.x * kSampleWarps + warp; r < rows; r += (long long)gridDim.x * kSampleWarps) {
    const int a = (int)(r % A);
    const float* x = ar.logits + r * P;
    float m = -3.402823466e38f;  // TF: max over the finite logits
    for (int k = lane; k < P; k += 32) {
      const float v = x[k];
      if (isfinite(v)) m = fmaxf(m, v);
    }
    m = warp_max(m);
    for (int k = lane; k < P; k += 32) {
      const float v = x[k];
      cdf[k] = isfinite(v) ? exp((double)v - (double)m) : 0.0;
    }
    __syncwarp();
    double total = 0.0;
    if (lane == 0) {  // sequential fp64 running total, exactly as the reference kernel accumulates it
      for (int k = 0; k < P; ++k) {
        total += cdf[k];
        cdf[k] = total;
      }
    }
    total = __shfl_sync(0xffffffffu, total, 0);
    __syncwarp();
    double u;
    if (ar.ext_uniform != nullptr) {
      u = ar.ext_uniform[r];
    } else {
      const uint4 q = rng(ar.offset, (uint64_t)r);
      u = u64_to_unit_double(q.x, q.y);
    }
    const double to_find = u * total;
    int first = P;  // upper_bound: first k with cdf[k] > to_find
    for (int k = lane; k < P; k += 32) {
      if (cdf[k] > to_find) {
        first = k;
        break;
      }
    }
    int idx = warp_min_i(first);
    // u*total == total can only happen through rounding; TF's upper_bound would return P there
    // (undefined class). Clamp to the last particle, which the oracle does as well.
    if (idx >= P) idx = P - 1;
    if (lane == 0) {
      float eps;
      if (ar.ext_normal != nullptr) {
        eps = ar.ext_normal[r * P + idx];
      } else {
        eps = philox_normal(rng, ar.offset + 1, (uint64_t)r);
      }
      const float mu = ar.loc[a * P + idx];
      const float sd = expf(ar.logstd[a * P + idx]);
      ar.action[r] = __fadd_rn(__fmul_rn(eps, sd), mu);  // Normal.sample: eps * scale + loc
      ar.idx[r] = idx;
    }
    __syncwarp();
  }
}
// K3 (reparameterised sample, forward + straight-through backward) lives in rsample.cu.

// deterministic action: plain -> loc[argmax logits]; tanh -> tanh(loc[argmax softmax(logits)])
__global__ void __launch_bounds__(kSampleWarps * 32) mean_kernel(const float* __restrict__ logits,
                                                                 const float* __restrict__ loc, float* __restrict__ action,
                                                                 int32_t* __restrict__ idx_out, int B, int A, int P,
                                                                 int tanh_flag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)B * A;
  for (long long r = (long long)blockIdx.x * kSampleWarps + warp; r < rows; r += (long long)gridDim.x * kSampleWarps) {
    const int a = (int)(r % A);
    const float* x = logits + r * P;
    float m = -3.402823466e38f;
    for (int k = lane; k < P; k += 32) m = fmaxf(m, x[k]);
    m = warp_max(m);
    int arg = 0x7fffffff;
    for (int k = lane; k < P; k += 32)
      if (x[k] == m) {
        arg = k;
        break;
      }
    arg = warp_min_i(arg);
    if (arg >= P) arg = 0;  // all-NaN row
    if (lane == 0) {
      const float mu = loc[a * P + arg];
      action[r] = tanh_flag ? tanhf(mu) : mu;
      if (idx_out) idx_out[r] = arg;
    }
  }
}

static int rows_grid(long long rows) {
  long long want = (rows + kSampleWarps - 1) / kSampleWarps;
  const long long cap = 148LL * 8;
  if (want > cap) want = cap;
  return (int)(want < 1 ? 1 : want);
}

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_head_sample(const pfpn_sample_args* args, pfpn_stream_t stream_) {
  if (!args) return PFPN_ERR_ARG;
  const pfpn_sample_args& a = *args;
  if (a.B < 0 || a.A <= 0 || a.P <= 0) return PFPN_ERR_ARG;
  if (a.B == 0) return PFPN_OK;
  if (!a.logits || !a.loc || !a.logstd || !a.action || !a.idx) return PFPN_ERR_ARG;
  const size_t smem = (size_t)kSampleWarps * a.P * sizeof(double);
  if (smem > 200 * 1024) return PFPN_ERR_UNSUPPORTED;
  PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_kernel<<<rows_grid((long long)a.B * a.A), kSampleWarps * 32, smem, reinterpret_cast<cudaStream_t>(stream_)>>>(a);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}


extern "C" int pfpn_head_mean(const float* logits, const float* loc, float* action, int32_t* idx, int32_t B, int32_t A,
                              int32_t P, uint32_t flags, pfpn_stream_t stream_) {
  if (!logits || !loc || !action || B < 0 || A <= 0 || P <= 0) return PFPN_ERR_ARG;
  if (B == 0) return PFPN_OK;
  mean_kernel<<<rows_grid((long long)B * A), kSampleWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      logits, loc, action, idx, B, A, P, (flags & PFPN_HEAD_FLAG_TANH) ? 1 : 0);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}
