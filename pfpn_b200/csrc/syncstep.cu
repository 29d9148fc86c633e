// The data-parallel optimizer step as THREE launches with no host-side state in their arguments, so that the whole DPPO
// minibatch update can be captured once in a CUDA graph and replayed:
//
//   sync_sumsq_kernel        partial sums of squares of this rank's gradients (fp64, fixed order) + the step counters
//   sync_clip_stage_kernel   global norm -> TF clip scale -> scaled gradients and the four pushed statistics written
//                            straight into this rank's peer-visible staging slot (parity = call & 1), then the LAST CTA
//                            publishes the call number in every peer's flag word (release, system scope)
//   sync_reduce_adam_kernel  waits for the N flags, sums the N staged buckets in rank order out of NVLink peer memory,
//                            mean, Adam (lr_t from the DEVICE step counter), averaged statistics assigned in place
//                            (one-phase: every rank reads every bucket;  two-phase: reduce-scatter + all-gather)
//
// Replaces (semantics): clip_grads -> tf.clip_by_global_norm BEFORE aggregation (models/workers/base_worker.py:97-102),
// SyncReplicasOptimizer's accumulator mean of gradients and pushed statistics (models/sync_model.py:37-45,92-96),
// ApplyAdam (base_worker.py:64-70).  Round 1 spent ~15 launches (incl. an 8.4 MB device-to-device copy and ~8 tiny torch
// copies) on this chain and needed the step number / buffer parity from the host on every call.
#include <string.h>

#include "common.cuh"

namespace pfpn {

constexpr int kSyncBlocks = 296;
constexpr int kSyncMaxPeers = 8;
constexpr long long kSyncSpinTimeout = 20000000000LL;

__device__ __forceinline__ void sync_st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int sync_ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sync_spin_until_ge(const int* f, int value) {
  const long long t0 = clock64();
  while (sync_ld_acquire_sys(f) < value) {
    if (clock64() - t0 > kSyncSpinTimeout) __trap();
  }
}

// counters (device int32[4]): [0] exchange calls on the peer buffers, [1] Adam step, [2] network global_step, [3] ticket
struct SyncK {
  float* grads;  // this rank's bucket [n_total]; gradients [0, n_params) are scaled IN PLACE when nranks == 1
  size_t n_params, n_total;
  float clip;
  const float* new_mean;
  const float* new_std;
  float* state_mean;
  float* state_std;
  float* max_active;
  float* sum_active;
  int S, AP;
  float* params;
  float* params_lo;
  float* m;
  float* v;
  float lr, b1, b2, eps;
  int* counters;
  float* norm_scale;
  double* part;
  float* stage[kSyncMaxPeers];    // rank r's staging base: [2][n_total]
  float* reduced[kSyncMaxPeers];  // rank r's averaged-slice buffer [n_total] (two-phase)
  int* flags[kSyncMaxPeers];      // rank r's flag words: [0,8) bucket published, [8,16) slice published, [32] ticket
  int rank, nranks;
  size_t slice4;
};

__global__ void __launch_bounds__(256) sync_sumsq_kernel(const SyncK k) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // nobody in this launch reads them; later launches see the new values
    k.counters[0] += 1;
    k.counters[1] += 1;
    k.counters[2] += 1;
  }
  __shared__ double sh[8];
  const float4* g4 = reinterpret_cast<const float4*>(k.grads);
  const size_t n4 = k.n_params / 4;
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4 x = g4[i];
    s += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    k.part[blockIdx.x] = t;
  }
}

// statistics tail layout: [new_mean S][new_std S][max_active AP][sum_active AP] (the normaliser part absent when S == 0)
__device__ __forceinline__ float sync_stat_src(const SyncK& k, int j) {
  if (j < k.S) return k.new_mean[j];
  if (j < 2 * k.S) return k.new_std[j - k.S];
  j -= 2 * k.S;
  return j < k.AP ? k.max_active[j] : k.sum_active[j - k.AP];
}
__device__ __forceinline__ void sync_stat_dst(const SyncK& k, int j, float x) {
  if (j < k.S) {
    k.state_mean[j] = x;
    return;
  }
  if (j < 2 * k.S) {
    k.state_std[j - k.S] = x;
    return;
  }
  j -= 2 * k.S;
  if (j < k.AP) k.max_active[j] = x;
  else k.sum_active[j - k.AP] = x;
}

__global__ void __launch_bounds__(256) sync_clip_stage_kernel(const SyncK k) {
  __shared__ float s_scale;
  __shared__ int s_last;
  // every CTA derives the norm itself from the partials, in ONE fixed order (lane-strided sums, xor tree)
  if (threadIdx.x < 32) {
    double t = 0.0;
    for (int i = threadIdx.x; i < kSyncBlocks; i += 32) t += k.part[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(t);
      float scale = 1.f;  // TF: clip * min(1/norm, 1/clip); NaN when the norm is not finite
      if (k.clip > 0.f) scale = isfinite(norm) ? k.clip * fminf(1.f / norm, 1.f / k.clip) : __int_as_float(0x7fc00000);
      s_scale = scale;
      if (blockIdx.x == 0) {
        k.norm_scale[0] = norm;
        k.norm_scale[1] = scale;
      }
    }
  }
  __syncthreads();
  const float scale = s_scale;
  const int value = k.counters[0];
  float* dst = k.nranks > 1 ? k.stage[k.rank] + (size_t)(value & 1) * k.n_total : k.grads;
  const size_t n4 = k.n_params / 4;
  const float4* g4 = reinterpret_cast<const float4*>(k.grads);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 x = g4[i];
    x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
    d4[i] = x;
  }
  const int n_stats = 2 * k.S + 2 * k.AP;
  for (int j = blockIdx.x * 256 + threadIdx.x; j < n_stats; j += gridDim.x * 256) dst[k.n_params + j] = sync_stat_src(k, j);
  if (k.nranks > 1) {
    __threadfence_system();  // my part of the staged bucket is visible system-wide before my ticket
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&k.counters[3], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x < k.nranks) {
      if (threadIdx.x == 0) k.counters[3] = 0;
      __threadfence_system();
      sync_st_release_sys(k.flags[threadIdx.x] + k.rank, value);
    }
  }
}

__device__ __forceinline__ float sync_lr_t(const SyncK& k) {
  const double t = (double)k.counters[1];
  return (float)((double)k.lr * sqrt(1.0 - pow((double)k.b2, t)) / (1.0 - pow((double)k.b1, t)));
}

#define PFPN_SYNC_ADAM4(s)                                                                                          \
  {                                                                                                                 \
    float4 mi = reinterpret_cast<float4*>(k.m)[i], vi = reinterpret_cast<float4*>(k.v)[i],                           \
           pi = reinterpret_cast<float4*>(k.params)[i];                                                             \
    mi.x = b1 * mi.x + (1.f - b1) * s.x; vi.x = b2 * vi.x + (1.f - b2) * s.x * s.x; pi.x -= lr_t * mi.x / (sqrtf(vi.x) + eps); \
    mi.y = b1 * mi.y + (1.f - b1) * s.y; vi.y = b2 * vi.y + (1.f - b2) * s.y * s.y; pi.y -= lr_t * mi.y / (sqrtf(vi.y) + eps); \
    mi.z = b1 * mi.z + (1.f - b1) * s.z; vi.z = b2 * vi.z + (1.f - b2) * s.z * s.z; pi.z -= lr_t * mi.z / (sqrtf(vi.z) + eps); \
    mi.w = b1 * mi.w + (1.f - b1) * s.w; vi.w = b2 * vi.w + (1.f - b2) * s.w * s.w; pi.w -= lr_t * mi.w / (sqrtf(vi.w) + eps); \
    reinterpret_cast<float4*>(k.m)[i] = mi;                                                                         \
    reinterpret_cast<float4*>(k.v)[i] = vi;                                                                         \
    reinterpret_cast<float4*>(k.params)[i] = pi;                                                                    \
    if (k.params_lo != nullptr) { /* the part of the new weights a tf32 operand drops: the GEMMs' B_lo operand */     \
      float4 lo;                                                                                                    \
      lo.x = pi.x - __uint_as_float(__float_as_uint(pi.x) & 0xffffe000u);                                           \
      lo.y = pi.y - __uint_as_float(__float_as_uint(pi.y) & 0xffffe000u);                                           \
      lo.z = pi.z - __uint_as_float(__float_as_uint(pi.z) & 0xffffe000u);                                           \
      lo.w = pi.w - __uint_as_float(__float_as_uint(pi.w) & 0xffffe000u);                                           \
      reinterpret_cast<float4*>(k.params_lo)[i] = lo;                                                               \
    }                                                                                                               \
  }

// One-phase: every rank reads every staged bucket (N-1 bucket volumes over NVLink per GPU); also the N == 1 path.
__global__ void __launch_bounds__(256) sync_reduce_adam_kernel(const SyncK k) {
  __shared__ float s_lr;
  const int value = k.counters[0];
  if (threadIdx.x == 0) s_lr = sync_lr_t(k);
  if (k.nranks > 1 && threadIdx.x < k.nranks) sync_spin_until_ge(k.flags[k.rank] + threadIdx.x, value);
  __syncthreads();
  const float lr_t = s_lr, b1 = k.b1, b2 = k.b2, eps = k.eps;
  const float inv_n = 1.f / (float)k.nranks;
  const size_t off = k.nranks > 1 ? (size_t)(value & 1) * k.n_total : 0;
  const float* bucket[kSyncMaxPeers];
#pragma unroll
  for (int p = 0; p < kSyncMaxPeers; ++p) bucket[p] = p < k.nranks ? (k.nranks > 1 ? k.stage[p] + off : k.grads) : nullptr;
  const size_t n4 = k.n_params / 4;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kSyncMaxPeers; ++p) {  // fixed order: identical result on every rank
      if (p < k.nranks) {
        const float4 g = __ldcg(reinterpret_cast<const float4*>(bucket[p]) + i);
        s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      }
    }
    s.x *= inv_n; s.y *= inv_n; s.z *= inv_n; s.w *= inv_n;
    PFPN_SYNC_ADAM4(s)
  }
  // the four pushed statistics: accumulator mean, assigned to the (replicated) variables
  const int n_stats = 2 * k.S + 2 * k.AP;
  for (int j = blockIdx.x * 256 + threadIdx.x; j < n_stats; j += gridDim.x * 256) {
    float s = 0.f;
    for (int p = 0; p < k.nranks; ++p) s += __ldcg(bucket[p] + k.n_params + j);
    sync_stat_dst(k, j, s * inv_n);
  }
}

// Two-phase (N >= 3): rank r averages only its 1/N slice of the bucket (phase 1), publishes it in its peer-visible
// `reduced` buffer, every rank gathers the N averaged slices and applies Adam (phase 2): 2 (N-1)/N bucket volumes per
// GPU.  Phase 2 waits for the local phase 1 of every CTA: the grid is launched cooperatively (co-residency guaranteed).
__global__ void __launch_bounds__(256) sync_reduce_adam_rs_kernel(const SyncK k) {
  __shared__ float s_lr;
  const int value = k.counters[0];
  const int rank = k.rank, nranks = k.nranks;
  if (threadIdx.x == 0) s_lr = sync_lr_t(k);
  if (threadIdx.x < nranks) sync_spin_until_ge(k.flags[rank] + threadIdx.x, value);
  __syncthreads();
  const float lr_t = s_lr, b1 = k.b1, b2 = k.b2, eps = k.eps;
  const float inv_n = 1.f / (float)nranks;
  const size_t off = (size_t)(value & 1) * k.n_total;
  const size_t n4 = k.n_total / 4;
  const size_t gtid = (size_t)blockIdx.x * 256 + threadIdx.x, gsz = (size_t)gridDim.x * 256;
  const size_t lo = (size_t)rank * k.slice4, hi = min(n4, lo + k.slice4);
  for (size_t i = lo + gtid; i < hi; i += gsz) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < nranks; ++p) {
      const float4 g = __ldcg(reinterpret_cast<const float4*>(k.stage[p] + off) + i);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    s.x *= inv_n; s.y *= inv_n; s.z *= inv_n; s.w *= inv_n;
    reinterpret_cast<float4*>(k.reduced[rank])[i] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    int* ctr = k.flags[rank] + 32;
    if (atomicAdd(ctr, 1) == (int)gridDim.x - 1) {  // last CTA of this call
      *ctr = 0;
      __threadfence_system();
      for (int p = 0; p < nranks; ++p) sync_st_release_sys(k.flags[p] + 8 + rank, value);
    }
  }
  const size_t np4 = k.n_params / 4;
  const int n_stats = 2 * k.S + 2 * k.AP;
  for (int kk = 0; kk < nranks; ++kk) {
    const int q = (rank + kk) % nranks;
    if (threadIdx.x == 0) sync_spin_until_ge(k.flags[rank] + 8 + q, value);
    __syncthreads();
    const size_t qlo = (size_t)q * k.slice4, qhi = min(n4, qlo + k.slice4);
    for (size_t i = qlo + gtid; i < qhi; i += gsz) {
      const float4 s = __ldcg(reinterpret_cast<const float4*>(k.reduced[q]) + i);
      if (i < np4) {
        PFPN_SYNC_ADAM4(s)
      } else {  // statistics tail (n_params % 4 == 0: a float4 never straddles the boundary)
        const int j = (int)(i * 4 - k.n_params);
        const float xs[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (j + c < n_stats) sync_stat_dst(k, j + c, xs[c]);
      }
    }
  }
}

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_sync_step_scratch_bytes(size_t* bytes) {
  if (!bytes) return PFPN_ERR_ARG;
  *bytes = kSyncBlocks * sizeof(double);
  return PFPN_OK;
}

extern "C" int pfpn_sync_step(const pfpn_sync_args* a, pfpn_stream_t stream_) {
  if (!a) return PFPN_ERR_ARG;
  if (!a->grads || !a->params || !a->m || !a->v || !a->counters || !a->norm_scale || !a->scratch) return PFPN_ERR_ARG;
  if (a->n_params == 0 || (a->n_params & 3) || (a->n_total & 3) || a->n_params > a->n_total) return PFPN_ERR_ARG;
  if (a->scratch_bytes < kSyncBlocks * sizeof(double)) return PFPN_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(a->scratch) & 7u) return PFPN_ERR_ALIGN;
  if (a->nranks < 1 || a->nranks > kSyncMaxPeers || a->rank < 0 || a->rank >= a->nranks) return PFPN_ERR_ARG;
  if (a->S < 0 || a->AP < 0) return PFPN_ERR_ARG;
  if (a->S > 0 && (!a->new_mean || !a->new_std || !a->state_mean || !a->state_std)) return PFPN_ERR_ARG;
  if (a->AP > 0 && (!a->max_active || !a->sum_active)) return PFPN_ERR_ARG;
  if (a->n_total < a->n_params + 2 * (size_t)a->S + 2 * (size_t)a->AP) return PFPN_ERR_ARG;
  SyncK k;
  memset(&k, 0, sizeof(k));
  k.grads = a->grads; k.n_params = a->n_params; k.n_total = a->n_total; k.clip = a->clip;
  k.new_mean = a->new_mean; k.new_std = a->new_std; k.state_mean = a->state_mean; k.state_std = a->state_std;
  k.max_active = a->max_active; k.sum_active = a->sum_active; k.S = a->S; k.AP = a->AP;
  k.params = a->params; k.params_lo = a->params_lo; k.m = a->m; k.v = a->v; k.lr = a->lr; k.b1 = a->beta1; k.b2 = a->beta2; k.eps = a->eps;
  k.counters = a->counters; k.norm_scale = a->norm_scale; k.part = reinterpret_cast<double*>(a->scratch);
  k.rank = a->rank; k.nranks = a->nranks;
  const bool two_phase = a->two_phase != 0 && a->nranks > 1;
  if (a->nranks > 1) {
    if (!a->stage || !a->flags || (two_phase && !a->reduced)) return PFPN_ERR_ARG;
    for (int p = 0; p < a->nranks; ++p) {
      if (!a->stage[p] || !a->flags[p] || (two_phase && !a->reduced[p])) return PFPN_ERR_ARG;
      k.stage[p] = a->stage[p];
      k.flags[p] = a->flags[p];
      k.reduced[p] = two_phase ? a->reduced[p] : nullptr;
    }
  }
  const size_t n4 = a->n_total / 4;
  k.slice4 = (n4 + a->nranks - 1) / a->nranks;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  sync_sumsq_kernel<<<kSyncBlocks, 256, 0, st>>>(k);
  PFPN_CUDA_OK(cudaGetLastError());
  sync_clip_stage_kernel<<<kSyncBlocks, 256, 0, st>>>(k);
  PFPN_CUDA_OK(cudaGetLastError());
  if (!two_phase) {
    sync_reduce_adam_kernel<<<kSyncBlocks, 256, 0, st>>>(k);
    PFPN_CUDA_OK(cudaGetLastError());
    return PFPN_OK;
  }
  int dev = 0, sms = 0, per_sm = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  PFPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PFPN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sync_reduce_adam_rs_kernel, 256, 0));
  if (per_sm < 1) return PFPN_ERR_UNSUPPORTED;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sms * (per_sm < 2 ? per_sm : 2)));
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  PFPN_CUDA_OK(cudaLaunchKernelEx(&cfg, sync_reduce_adam_rs_kernel, k));
  return PFPN_OK;
}
