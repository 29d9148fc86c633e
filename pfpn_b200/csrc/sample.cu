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
__device__ __forceinline__ float philox_normal(const Philox7& rng, uint64_t offset, uint64_t ctr) {
  const uint4 r = rng(offset, ctr);
  return normal_from_bits(r.x, r.y);
}

// ------------------------------------------------------------------------------------------------
// K2: plain sample
// ------------------------------------------------------------------------------------------------
// TF's CPU Multinomial functor works in fp64 (exp of every logit, sequential running total, upper_bound), and fp64
// transcendental throughput is what a literal port spends its time on.  The selected index only depends on which
// CDF interval u*total falls into, so the kernel finds the interval with fp32 arithmetic whose error is bounded
// (|cdf32_k - cdf_k| <= delta * total) and accepts it only when u*total keeps a 2*delta*total margin to both
// interval ends; otherwise (~0.1 % of rows) that row is redone literally in fp64.  Indices are therefore exactly
// those of the fp64 algorithm.
//
// A warp walks a chunk of up to 32 consecutive rows.  Phase 1 (lanes over particles, coalesced): fp32 terms
// 2^((l_k - max) log2e) to shared memory.  Phase 2 (lane j owns row j): sequential running total, binary-search
// upper_bound, margin test / fp64 fallback, the location draw for the chosen particle only, coalesced outputs.
__device__ __forceinline__ float exp_term_f32(float d) {  // e^d for d <= 0, relative error ~5e-7
  const float t = d * kLog2e;
  const float n = rintf(t);
  float f = fmaf(d, kLog2e, -n);             // single rounding of d*log2e - n
  f = fmaf(d, 1.925963033e-8f, f);           // low part of log2(e) beyond its fp32 value
  const int ni = (int)n;
  const float scale = ni >= -126 ? __int_as_float((ni + 127) << 23) : 0.f;
  return ex2f(f) * scale;
}

// the literal fp64 algorithm for one row, by one thread (fallback; also the semantic definition)
__device__ __noinline__ int multinomial_row_fp64(const float* __restrict__ x, int P, double u) {
  float m = -3.402823466e38f;
  for (int k = 0; k < P; ++k) {
    const float v = x[k];
    if (isfinite(v)) m = fmaxf(m, v);
  }
  double total = 0.0;
  for (int k = 0; k < P; ++k) {
    const float v = x[k];
    if (isfinite(v)) total += exp((double)v - (double)m);
  }
  const double to_find = u * total;
  double run = 0.0;
  for (int k = 0; k < P; ++k) {  // first k with cdf[k] > to_find (== std::upper_bound on the running totals)
    const float v = x[k];
    if (isfinite(v)) run += exp((double)v - (double)m);
    if (to_find < run) return k;
  }
  // u*total == total can only happen through rounding; TF's upper_bound would return P there
  // (undefined class). Clamp to the last particle, which the oracle does as well.
  return P - 1;
}

__global__ void __launch_bounds__(kSampleWarps * 32) sample_kernel(const pfpn_sample_args ar, const int chunk_rows, const int PS) {
  extern __shared__ float cdf_s[];  // [kSampleWarps][chunk_rows][PS], PS = P | 1 (odd stride: conflict-free row-per-lane)
  const int A = ar.A, P = ar.P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* cdf = cdf_s + (size_t)warp * chunk_rows * PS;
  const long long rows = (long long)ar.B * A;
  const long long nchunks = (rows + chunk_rows - 1) / chunk_rows;
  const Philox7 rng(ar.seed);  // (seven rounds: the fewest that pass BigCrush; ten cost 30 % more integer work per draw)
  // error budget of the fp32 CDF relative to the total: per-term 2e-6 + (P - 1) 2^-24 from the sequential sum
  const float delta = 4e-6f + (float)P * 6e-8f;
  for (long long c = (long long)blockIdx.x * kSampleWarps + warp; c < nchunks; c += (long long)gridDim.x * kSampleWarps) {
    const long long r0 = c * chunk_rows;
    const int nrow = (int)min((long long)chunk_rows, rows - r0);
    for (int j = 0; j < nrow; ++j) {
      const float* x = ar.logits + (r0 + j) * P;
      float m = -3.402823466e38f;  // TF: max over the finite logits
      for (int k = lane; k < P; k += 32) {
        const float v = x[k];
        if (isfinite(v)) m = fmaxf(m, v);
      }
      m = warp_max(m);
      for (int k = lane; k < P; k += 32) {
        const float v = x[k];
        cdf[j * PS + k] = isfinite(v) ? exp_term_f32(v - m) : 0.f;
      }
    }
    __syncwarp();
    if (lane < nrow) {
      const long long r = r0 + lane;
      float* row = cdf + lane * PS;
      float total = 0.f;
      for (int k = 0; k < P; ++k) {
        total += row[k];
        row[k] = total;
      }
      double u;
      if (ar.ext_uniform != nullptr) {
        u = ar.ext_uniform[r];
      } else {
        const uint4 q = rng(ar.offset, (uint64_t)r);
        u = u64_to_unit_double(q.x, q.y);
      }
      const float to_find = (float)(u * (double)total);
      int lo = 0, len = P;  // first k with cdf32[k] > to_find
      while (len > 0) {
        const int half = len >> 1;
        if (!(to_find < row[lo + half])) {
          lo += half + 1;
          len -= half + 1;
        } else {
          len = half;
        }
      }
      // accept only when u*total is provably inside (cdf[lo-1], cdf[lo]) of the fp64 CDF as well
      const float margin = 2.f * delta * total + 1.2e-7f * total;  // + the fp32 rounding of to_find itself
      bool certain = lo < P && total > 0.f && isfinite(total) && (row[lo] - to_find > margin) &&
                     (lo == 0 || to_find - row[lo - 1] > margin);
      int idx = lo;
      if (!certain) idx = multinomial_row_fp64(ar.logits + r * P, P, u);
      const int a = (int)(r % A);
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
  const int PS = a.P | 1;
  const long long rows = (long long)a.B * a.A;
  // rows per warp chunk: 32 at scale; fewer when the rows would not fill the machine or shared memory (~12 KiB per warp)
  long long chunk = rows / (148LL * 4 * kSampleWarps);
  const long long smem_cap = (12 * 1024) / ((long long)PS * (long long)sizeof(float));
  if (chunk > 32) chunk = 32;
  if (chunk > smem_cap) chunk = smem_cap;
  if (chunk < 1) chunk = 1;
  const size_t smem = (size_t)kSampleWarps * chunk * PS * sizeof(float);
  if (smem > 200 * 1024) return PFPN_ERR_UNSUPPORTED;
  PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nchunks = (rows + chunk - 1) / chunk;
  sample_kernel<<<rows_grid(nchunks), kSampleWarps * 32, smem, reinterpret_cast<cudaStream_t>(stream_)>>>(a, (int)chunk, PS);
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
