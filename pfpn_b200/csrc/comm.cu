// K7 fused with the data-parallel exchange C1: one kernel that sums the clipped gradient buckets of
// all ranks straight out of peer memory (NVLink 5 / NVSwitch P2P loads) and applies the averaged
// gradient with Adam, instead of NCCL all-reduce + a separate Adam launch.
//
// Replaces (semantics): SyncReplicasOptimizer's accumulator mean + ApplyAdam
//   /root/reference/models/sync_model.py:92-96, models/workers/base_worker.py:64-70.
//
// Protocol (one process per GPU, buffers shared through CUDA IPC by the host side):
//   1. every rank writes its clipped [gradients | statistics] bucket into its own staging buffer
//      (double-buffered by step parity) and then runs `peer_signal_kernel`, which publishes the step
//      number into the flag array of every peer (system-scope fence + remote store);
//   2. `peer_allreduce_adam_kernel` spins until all flags carry the step number, then every rank
//      sums the N buckets in rank order 0..N-1 -- the same order on every rank, so the replicas stay
//      bit-identical without a broadcast -- scales by 1/N, updates params / m / v for the gradient
//      part and writes the averaged statistics back for the caller.
// A staging buffer of parity p is rewritten two steps later; by then every peer has signalled the
// step in between, which it does only after its own reduce kernel of the earlier step finished.
#include <string.h>

#include "common.cuh"

namespace pfpn {

constexpr int kMaxPeers = 8;
// A peer that never publishes (crashed rank, mismatched call counts) must not wedge the GPU: after ~10 s of polling
// the kernel traps, which surfaces as a launch failure on the host instead of a hang.
constexpr long long kSpinTimeoutCycles = 20000000000LL;
// Flags are written with st.release.sys and polled with ld.acquire.sys: the data a peer published before raising its flag
// (its bucket, written by earlier kernels on its stream or by the same kernel before a system-scope fence) is visible
// to whoever observed the flag.
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until_ge(const int* f, int value) {
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < value) {
    if (clock64() - t0 > kSpinTimeoutCycles) __trap();
  }
}
struct PeerPtrs {
  const float* bucket[kMaxPeers];
  int* flags[kMaxPeers];
};

__global__ void peer_signal_kernel(PeerPtrs pp, int rank, int nranks, int value) {
  const int p = threadIdx.x;
  if (p < nranks) {
    __threadfence_system();  // this rank's bucket (written by earlier kernels on the stream) -> visible
    st_release_sys(pp.flags[p] + rank, value);
  }
}

__global__ void __launch_bounds__(256) peer_allreduce_adam_kernel(PeerPtrs pp, int rank, int nranks, int value,
                                                                  size_t n_params, size_t n_total, float* __restrict__ params,
                                                                  float* __restrict__ m, float* __restrict__ v,
                                                                  float* __restrict__ avg_out, float lr_t, float b1, float b2,
                                                                  float eps) {
  if (threadIdx.x < nranks) spin_until_ge(pp.flags[rank] + threadIdx.x, value);
  __syncthreads();
  const float inv_n = 1.f / (float)nranks;
  const size_t n4 = n_total / 4;  // buckets are padded to a multiple of 4 floats
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < nranks; ++p) {  // fixed order: identical result on every rank
      const float4 g = __ldcg(reinterpret_cast<const float4*>(pp.bucket[p]) + i);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    s.x *= inv_n; s.y *= inv_n; s.z *= inv_n; s.w *= inv_n;
    reinterpret_cast<float4*>(avg_out)[i] = s;
    if (i * 4 < n_params) {  // n_params % 4 == 0
      float4 mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i], pi = reinterpret_cast<float4*>(params)[i];
#define PFPN_ADAM1(c)                                   \
  mi.c = b1 * mi.c + (1.f - b1) * s.c;                  \
  vi.c = b2 * vi.c + (1.f - b2) * s.c * s.c;            \
  pi.c -= lr_t * mi.c / (sqrtf(vi.c) + eps);
      PFPN_ADAM1(x) PFPN_ADAM1(y) PFPN_ADAM1(z) PFPN_ADAM1(w)
#undef PFPN_ADAM1
      reinterpret_cast<float4*>(m)[i] = mi;
      reinterpret_cast<float4*>(v)[i] = vi;
      reinterpret_cast<float4*>(params)[i] = pi;
    }
  }
}

// Two-phase variant for N >= 3 (reduce-scatter + all-gather inside one kernel): reading all N buckets on every
// rank moves (N-1) x 8.4 MB per GPU over NVLink; here rank r sums only its 1/N slice (phase 1), publishes the
// averaged slice in its own peer-visible `red` buffer, and every rank then gathers the N averaged slices and
// applies Adam to all parameters (phase 2): 2 (N-1)/N x 8.4 MB per GPU.  Flag words per rank: [0,8) bucket
// published (by peer), [8,16) averaged slice published (by peer), [32] local CTA-arrival counter.
// Replicas stay bit-identical: every element is summed once, in rank order, by its owning rank.
struct PeerPtrs2 {
  const float* bucket[kMaxPeers];
  float* red[kMaxPeers];
  int* flags[kMaxPeers];
};
__global__ void __launch_bounds__(256) peer_allreduce_adam_rs_kernel(PeerPtrs2 pp, int rank, int nranks, int value, size_t n_params,
                                                                     size_t n_total, size_t slice4, float* __restrict__ params,
                                                                     float* __restrict__ m, float* __restrict__ v,
                                                                     float* __restrict__ avg_out, float lr_t, float b1, float b2,
                                                                     float eps) {
  const size_t n4 = n_total / 4;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  // ---- phase 0: every peer has published its clipped bucket of this step
  if (threadIdx.x < nranks) spin_until_ge(pp.flags[rank] + threadIdx.x, value);
  __syncthreads();
  // ---- phase 1: mean of my slice, in rank order
  const float inv_n = 1.f / (float)nranks;
  const size_t lo = (size_t)rank * slice4, hi = min(n4, lo + slice4);
  for (size_t i = lo + gtid; i < hi; i += gsz) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < nranks; ++p) {
      const float4 g = __ldcg(reinterpret_cast<const float4*>(pp.bucket[p]) + i);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    s.x *= inv_n; s.y *= inv_n; s.z *= inv_n; s.w *= inv_n;
    reinterpret_cast<float4*>(pp.red[rank])[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    int* ctr = pp.flags[rank] + 32;
    const int arrived = atomicAdd(ctr, 1) + 1;
    if (arrived == (int)gridDim.x) {  // last CTA of this call: every CTA has arrived, nobody touches the counter again
      *ctr = 0;                       // self-resetting (the next call on this stream starts from zero)
      __threadfence_system();
      for (int p = 0; p < nranks; ++p) st_release_sys(pp.flags[p] + 8 + rank, value);
    }
  }
  // ---- phase 2: gather the averaged slices (own slice first, then the peers round-robin) + Adam on everything
  for (int k = 0; k < nranks; ++k) {
    const int q = (rank + k) % nranks;
    if (threadIdx.x == 0) spin_until_ge(pp.flags[rank] + 8 + q, value);
    __syncthreads();
    const size_t qlo = (size_t)q * slice4, qhi = min(n4, qlo + slice4);
    for (size_t i = qlo + gtid; i < qhi; i += gsz) {
      const float4 s = __ldcg(reinterpret_cast<const float4*>(pp.red[q]) + i);
      reinterpret_cast<float4*>(avg_out)[i] = s;
      if (i * 4 < n_params) {
        float4 mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i], pi = reinterpret_cast<float4*>(params)[i];
#define PFPN_ADAM1(c)                                   \
  mi.c = b1 * mi.c + (1.f - b1) * s.c;                  \
  vi.c = b2 * vi.c + (1.f - b2) * s.c * s.c;            \
  pi.c -= lr_t * mi.c / (sqrtf(vi.c) + eps);
        PFPN_ADAM1(x) PFPN_ADAM1(y) PFPN_ADAM1(z) PFPN_ADAM1(w)
#undef PFPN_ADAM1
        reinterpret_cast<float4*>(m)[i] = mi;
        reinterpret_cast<float4*>(v)[i] = vi;
        reinterpret_cast<float4*>(params)[i] = pi;
      }
    }
  }
}

// Small exchange of the head's [2, A, P] particle gradients (the only collective of the
// sharded head path): signal + wait + ordered sum in ONE kernel.  The bucket was written by the kernel
// before this one on the stream (head_finalize writes straight into the staging buffer).
__global__ void __launch_bounds__(256) peer_sum_kernel(PeerPtrs pp, int rank, int nranks, int value, size_t n4,
                                                       float* __restrict__ out, float scale) {
  if (blockIdx.x == 0 && threadIdx.x < nranks) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + rank, value);  // publish to every peer (and self)
  }
  if (threadIdx.x < nranks) spin_until_ge(pp.flags[rank] + threadIdx.x, value);
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < nranks; ++p) {  // fixed order: identical result on every rank
      const float4 g = __ldcg(reinterpret_cast<const float4*>(pp.bucket[p]) + i);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    reinterpret_cast<float4*>(out)[i] = s;
  }
}

// Consumer side of the PUSH protocol (pfpn_head_logprob_push): the peers' kernels stored their rows into THIS rank's
// gather buffer and raised this rank's flag words; wait for the N flags, then sum the N local rows in rank order.
// Launched with programmatic stream serialization: it becomes resident behind K1's finalize and is already polling
// when the flags arrive.  No CTA waits on another CTA of this grid, so there is no co-residency requirement.
__global__ void __launch_bounds__(64) peer_gather_sum_kernel(const float* __restrict__ gather, const int* __restrict__ flags,
                                                              int nranks, int value, size_t n4, float* __restrict__ out,
                                                              float scale) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // (orders the previous consumer of `out` / this rank's own push)
  if (threadIdx.x < nranks) spin_until_ge(flags + threadIdx.x, value);
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < nranks; ++p) {  // fixed order: identical result on every rank
      const float4 g = __ldcg(reinterpret_cast<const float4*>(gather) + (size_t)p * n4 + i);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    reinterpret_cast<float4*>(out)[i] = s;
  }
}

// Consumer of a packet exchange (pfpn_head_push protocol 1): element i of every row carries its own sequence word.
__global__ void __launch_bounds__(64) peer_gather_sum_packets_kernel(const uint2* __restrict__ rows, int nranks, int value, size_t n,
                                                                      float* __restrict__ out, float scale) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // (orders the previous consumer of `out`)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint2 pk[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < nranks) asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(pk[r].x), "=r"(pk[r].y) : "l"(rows + r * n + i) : "memory");
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) {
      if (r < nranks) {
        while ((int)pk[r].y != value)
          asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(pk[r].x), "=r"(pk[r].y) : "l"(rows + r * n + i) : "memory");
        s += __uint_as_float(pk[r].x);  // fixed order: identical result on every rank
      }
    }
    out[i] = s * scale;
  }
}

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_peer_gather_sum_packets(const void* rows, int32_t nranks, int32_t value, size_t n, float* out, float scale,
                                            pfpn_stream_t stream_) {
  if (!rows || !out || nranks < 1 || nranks > kMaxPeers || n == 0 || value < 1 || (reinterpret_cast<uintptr_t>(rows) & 7))
    return PFPN_ERR_ARG;
  size_t grid = (n + 63) / 64;
  if (grid > 296) grid = 296;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(64);
  cfg.stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  PFPN_CUDA_OK(cudaLaunchKernelEx(&cfg, peer_gather_sum_packets_kernel, reinterpret_cast<const uint2*>(rows), (int)nranks, (int)value, n,
                                  out, scale));
  return PFPN_OK;
}

extern "C" int pfpn_peer_gather_sum(const float* gather, const int32_t* flags, int32_t nranks, int32_t value, size_t n,
                                    float* out, float scale, pfpn_stream_t stream_) {
  if (!gather || !flags || !out || nranks < 1 || nranks > kMaxPeers || (n & 3) || n == 0 || value < 1) return PFPN_ERR_ARG;
  // 64-thread CTAs (~2 K registers each): small enough to become resident NEXT TO the persistent head kernel's CTAs (which
  // leave 4 K registers and 28 KB of shared memory per SM), so a consumer on a second stream overlaps the next step
  const size_t n4 = n / 4;
  size_t grid = (n4 + 63) / 64;
  if (grid > 296) grid = 296;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(64);
  cfg.stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  PFPN_CUDA_OK(cudaLaunchKernelEx(&cfg, peer_gather_sum_kernel, gather, (const int*)flags, (int)nranks, (int)value, n4, out, scale));
  return PFPN_OK;
}

extern "C" int pfpn_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  if (dev == peer_device) return PFPN_OK;
  int can = 0;
  PFPN_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) return PFPN_ERR_UNSUPPORTED;
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return PFPN_OK;
  }
  return (int)e;
}

static int fill_peers(PeerPtrs* pp, const float* const* buckets, int32_t* const* flags, int nranks) {
  if (!buckets || !flags || nranks < 1 || nranks > kMaxPeers) return PFPN_ERR_ARG;
  for (int p = 0; p < kMaxPeers; ++p) {
    pp->bucket[p] = p < nranks ? buckets[p] : nullptr;
    pp->flags[p] = p < nranks ? flags[p] : nullptr;
    if (p < nranks && (!buckets[p] || !flags[p])) return PFPN_ERR_ARG;
  }
  return PFPN_OK;
}

// `buckets` / `flags`: HOST arrays of nranks device pointers (peer-mapped), indexed by rank.
extern "C" int pfpn_peer_signal(const float* const* buckets, int32_t* const* flags, int32_t rank, int32_t nranks,
                                int32_t value, pfpn_stream_t stream_) {
  PeerPtrs pp;
  int rc = fill_peers(&pp, buckets, flags, nranks);
  if (rc != PFPN_OK) return rc;
  peer_signal_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(pp, rank, nranks, value);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_peer_allreduce_adam(const float* const* buckets, int32_t* const* flags, int32_t rank, int32_t nranks,
                                        int32_t value, size_t n_params, size_t n_total, float* params, float* m, float* v,
                                        float* avg_out, float lr, float beta1, float beta2, float eps, int64_t step,
                                        pfpn_stream_t stream_) {
  PeerPtrs pp;
  int rc = fill_peers(&pp, buckets, flags, nranks);
  if (rc != PFPN_OK) return rc;
  if (!params || !m || !v || !avg_out || (n_params & 3) || (n_total & 3) || n_params > n_total || step < 1) return PFPN_ERR_ARG;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  peer_allreduce_adam_kernel<<<296, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      pp, rank, nranks, value, n_params, n_total, params, m, v, avg_out, (float)lr_t, beta1, beta2, eps);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

// Two-phase (reduce-scatter + all-gather) form of pfpn_peer_allreduce_adam for N >= 3.  `reduced`: HOST array of nranks
// peer-mapped device buffers of n_total floats (one per rank, written only by its owner); `flags` buffers must hold >= 64
// ints, zero-initialised; `value` must be 1, 2, 3, ... on successive calls (it is also the call index of the CTA counter).
extern "C" int pfpn_peer_allreduce_adam_rs(const float* const* buckets, float* const* reduced, int32_t* const* flags, int32_t rank,
                                           int32_t nranks, int32_t value, size_t n_params, size_t n_total, float* params, float* m,
                                           float* v, float* avg_out, float lr, float beta1, float beta2, float eps, int64_t step,
                                           pfpn_stream_t stream_) {
  PeerPtrs pp1;
  int rc = fill_peers(&pp1, buckets, flags, nranks);
  if (rc != PFPN_OK) return rc;
  if (!reduced || !params || !m || !v || !avg_out || (n_params & 3) || (n_total & 3) || n_params > n_total || step < 1 || value < 1)
    return PFPN_ERR_ARG;
  PeerPtrs2 pp;
  for (int p = 0; p < kMaxPeers; ++p) {
    pp.bucket[p] = pp1.bucket[p];
    pp.flags[p] = pp1.flags[p];
    pp.red[p] = p < nranks ? reduced[p] : nullptr;
    if (p < nranks && !reduced[p]) return PFPN_ERR_ARG;
  }
  const size_t n4 = n_total / 4;
  const size_t slice4 = (n4 + nranks - 1) / nranks;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  // Phase 2 waits for the LOCAL phase 1 of every CTA of this grid: the grid must be co-resident.  A cooperative launch
  // makes the driver guarantee that (or fail the launch) instead of assuming an idle GPU; the grid is sized from the
  // occupancy of this device.
  int dev = 0, sms = 0, per_sm = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  PFPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PFPN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer_allreduce_adam_rs_kernel, 256, 0));
  if (per_sm < 1) return PFPN_ERR_UNSUPPORTED;
  const int grid = sms * (per_sm < 2 ? per_sm : 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(256);
  cfg.stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  PFPN_CUDA_OK(cudaLaunchKernelEx(&cfg, peer_allreduce_adam_rs_kernel, pp, (int)rank, (int)nranks, (int)value, n_params, n_total,
                                  slice4, params, m, v, avg_out, (float)lr_t, beta1, beta2, eps));
  return PFPN_OK;
}

// out[n] = scale * sum over ranks (rank order) of buckets[r][n]; n % 4 == 0, n small (one wave of CTAs: every CTA
// spins on the local flags, so the grid must be co-resident).  `value` must increase by one per call.
extern "C" int pfpn_peer_allreduce_sum(const float* const* buckets, int32_t* const* flags, int32_t rank, int32_t nranks,
                                       int32_t value, size_t n, float* out, float scale, pfpn_stream_t stream_) {
  PeerPtrs pp;
  int rc = fill_peers(&pp, buckets, flags, nranks);
  if (rc != PFPN_OK) return rc;
  if (!out || (n & 3) || n == 0 || rank < 0 || rank >= nranks) return PFPN_ERR_ARG;
  const size_t n4 = n / 4;
  size_t grid = (n4 + 255) / 256;
  if (grid > 148) grid = 148;
  peer_sum_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(pp, rank, nranks, value, n4, out, scale);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

// ---- peer-visible allocations (cudaMalloc + CUDA IPC), opened on the CALLER's current device ----------
extern "C" int pfpn_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return PFPN_ERR_ARG;
  PFPN_CUDA_OK(cudaMalloc(ptr, bytes));
  PFPN_CUDA_OK(cudaMemset(*ptr, 0, bytes));
  PFPN_CUDA_OK(cudaDeviceSynchronize());
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  PFPN_CUDA_OK(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return PFPN_OK;
}
extern "C" int pfpn_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return PFPN_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  PFPN_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PFPN_OK;
}
extern "C" int pfpn_peer_close(void* ptr) {
  if (!ptr) return PFPN_ERR_ARG;
  PFPN_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return PFPN_OK;
}
extern "C" int pfpn_peer_free(void* ptr) {
  if (!ptr) return PFPN_ERR_ARG;
  PFPN_CUDA_OK(cudaFree(ptr));
  return PFPN_OK;
}
