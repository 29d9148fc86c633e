// K3 -- reparameterised sampling (SAC, normalize_output=True): forward and straight-through backward.
//
// Reference: /root/reference/networks/utils.py:156-186 -- w = softmax(logits + Gumbel(U)) (TFP 0.7
// RelaxedOneHotCategorical, T = 1), k* = argmax w, s_ = (loc + scale*eps)[k*], sample = tanh(s_),
// custom gradients `mask2` (:164-171) and `mask` (:176-183); closed form in SURVEY.md Appendix A2.
//
// A warp owns a chunk of up to 32 consecutive mixture rows (b, a) and walks them one at a time; lane l
// holds particles l, l+32, ... of the current row in registers (MAXE slots).  Whatever is needed once
// per row (the winner's location draw, tanh, the dloc / dlogstd scatter, the output stores) is deferred
// to a lane-parallel tail: lane j finishes row j of the chunk, so that work costs 1/32 per row.
// The draws are never stored: the backward regenerates them from the same Philox counters.
// Two arithmetic modes:
//   * verification (ext_uniform / ext_normal supplied): accurate libm functions, so that the argmax
//     particle is bit-exact against the oracle fed with the same draws;
//   * production (Philox): two counter streams -- one Philox4x32-7 block gives the Gumbel uniforms
//     of FOUR particles, one more gives their four N(0,1) draws (two Box-Muller pairs).  The forward
//     only regenerates the winner's normal block; the backward needs tanh(p_k) of every particle and
//     draws them all.  Transcendental work is on the MUFU pipe (lg2 / ex2 / sin / cos / rsq / rcp).
//     The kernel is RNG/ALU bound, not HBM bound, so this is where the time goes.
#include "common.cuh"

namespace pfpn {

constexpr int kRsWarps = 8;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_max_f32(float v) {  // CREDUX.MAX.F32 (sm_100a)
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}
__device__ __forceinline__ float selp_f32(bool c, float a, float b) {
  float r;
  asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f32 %0, %1, %2, p; }" : "=f"(r) : "f"(a), "f"(b), "r"((int)c));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_tanh(float x) {  // 1 - 2 / (1 + e^{2x}); |err| ~ 1e-7
  const float e = ex2f(x * (2.f * kLog2e));
  return 1.f - 2.f * rcpf(1.f + e);
}
// 23 random bits -> (i + 0.5) 2^-23, strictly inside (0, 1), without an int->float conversion
__device__ __forceinline__ float bits_to_unit(uint32_t w) {
  return __uint_as_float(0x3f800000u | (w >> 9)) - 0.99999994f;
}

constexpr uint64_t kNormalStream = 1ull << 63;  // counter bit separating the eps draws from the Gumbel draws
// four N(0,1) draws from one Philox block: two Box-Muller pairs on the MUFU pipe
__device__ __forceinline__ void normal4(const uint4 q, float (&n)[4]) {
  const float r0 = sqrt_approx(-2.f * kLn2 * lg2f(bits_to_unit(q.x)));
  const float r1 = sqrt_approx(-2.f * kLn2 * lg2f(bits_to_unit(q.z)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * bits_to_unit(q.y), &s0, &c0);
  __sincosf(6.283185307179586f * bits_to_unit(q.w), &s1, &c1);
  n[0] = r0 * c0;
  n[1] = r0 * s0;
  n[2] = r1 * c1;
  n[3] = r1 * s1;
}

// FAST keeps the noisy logits in the log2 domain and drops the constant -ln(ln 2) that the two-log
// Gumbel form carries: argmax and softmax are shift invariant.
template <bool BWD, bool FAST, int MAXE>
__global__ void __launch_bounds__(BWD ? 384 : kRsWarps * 32) rsample_kernel(const pfpn_rsample_args ar, const int chunk_rows,
                                                                            float* __restrict__ part) {
  // Backward mapping: a CTA walks whole states (A consecutive rows, contiguous in memory) and warp w owns the action
  // dimensions a = w, w + nw, ... of every state it sees.  The straight-through gradients of the winning particles
  // therefore reach acc[a][k*] from ONE warp, in state order, as plain read-modify-writes by the lane that owns k* --
  // no atomics -- and the per-CTA tables are combined in CTA order by rsample_bwd_finalize_kernel: bit-reproducible.
  extern __shared__ float smem_f[];  // BWD: acc[2][A*P] per-CTA dloc / dlogstd partials, then loc[A*P], sd[A*P]
  const int A = ar.A, P = ar.P, AP = A * P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* acc_s = smem_f;
  float* loc_s = smem_f + 2 * AP;
  float* sd_s = smem_f + 3 * AP;
  if (BWD) {
    for (int i = threadIdx.x; i < AP; i += blockDim.x) {
      acc_s[i] = 0.f;
      acc_s[AP + i] = 0.f;
      loc_s[i] = ar.loc[i];
      sd_s[i] = FAST ? ex2f(ar.logstd[i] * kLog2e) : expf(ar.logstd[i]);
    }
    __syncthreads();
  }
  const long long rows = (long long)ar.B * A;
  const long long nchunks = (rows + chunk_rows - 1) / chunk_rows;
  const int nw = blockDim.x >> 5;
  const Philox7 rng(ar.seed);
  const uint64_t rng_off = ar.offset + (ar.offset_dev != nullptr ? *ar.offset_dev : 0ull);  // (device word: graph replays)
  // forward: c = chunk of consecutive rows per warp; backward: c = state per CTA, its rows split over the warps by a
  const long long c_begin = BWD ? (long long)blockIdx.x : (long long)blockIdx.x * nw + warp;
  const long long c_step = BWD ? (long long)gridDim.x : (long long)gridDim.x * nw;
  const long long c_end = BWD ? (long long)ar.B : nchunks;
  for (long long c = c_begin; c < c_end; c += c_step) {
    const long long r0 = BWD ? c * A : c * chunk_rows;
    const int nrow = BWD ? A : (int)min((long long)chunk_rows, rows - r0);
    const int a0 = BWD ? 0 : (int)(r0 % A);
    int a = BWD ? warp : a0;
    int my_arg = 0;  // forward: per-row result kept lane-parallel (lane j <-> row r0 + j) for the tail
    for (int j = BWD ? warp : 0; j < nrow; j += BWD ? nw : 1) {
      const long long r = r0 + j;
      const float* x = ar.logits + r * P;
      float y[MAXE], pk[MAXE], ek[MAXE];
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        y[e] = -3.402823466e38f;
        pk[e] = 0.f;
        ek[e] = 0.f;
      }
      if (FAST) {
        // Gumbel stream: one Philox block -> the uniforms of FOUR particles of this lane
#pragma unroll
        for (int e4 = 0; e4 < MAXE; e4 += 4) {
          const int k0 = lane + 32 * e4;
          if (k0 < P) {
            const uint4 q = rng(rng_off, (uint64_t)(r * P + k0));
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (e4 + i < MAXE && k0 + 32 * i < P) {
                const float nl = fmaxf(-lg2f(bits_to_unit(w4[i])), 4e-8f);  // -log2 u, kept off 0 (u -> 1)
                y[e4 + i] = fmaf(x[k0 + 32 * i], kLog2e, -lg2f(nl));        // log2e * (logits + Gumbel) + const
              }
            }
          }
        }
        if (BWD) {
          // normal stream: one block -> two Box-Muller pairs -> the eps of the same four particles
#pragma unroll
          for (int e4 = 0; e4 < MAXE; e4 += 4) {
            const int k0 = lane + 32 * e4;
            if (k0 < P) {
              float n4[4];
              normal4(rng(rng_off, kNormalStream | (uint64_t)(r * P + k0)), n4);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int k = k0 + 32 * i;
                if (e4 + i < MAXE && k < P) {
                  ek[e4 + i] = n4[i];
                  pk[e4 + i] = fmaf(n4[i], sd_s[a * P + k], loc_s[a * P + k]);
                }
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
          const int k = lane + 32 * e;
          if (k < P) {
            y[e] = x[k] + (-logf(-logf(ar.ext_uniform[r * P + k])));  // (G + logits) / T, T = 1
            if (BWD) {
              ek[e] = ar.ext_normal[r * P + k];
              pk[e] = __fadd_rn(__fmul_rn(ek[e], sd_s[a * P + k]), loc_s[a * P + k]);
            }
          }
        }
      }
      float m = y[0];
#pragma unroll
      for (int e = 1; e < MAXE; ++e) m = fmaxf(m, y[e]);
      m = warp_max_f32(m);
      // argmax of w = argmax of the noisy logits; ties -> smallest index (tf.argmax)
      int arg = 0;
#pragma unroll
      for (int e = MAXE - 1; e >= 0; --e) {
        const unsigned hit = __ballot_sync(kFull, lane + 32 * e < P && y[e] == m);
        if (hit) arg = 32 * e + __ffs(hit) - 1;
      }
      if (!BWD) {
        if (lane == j) my_arg = arg;
      } else {
        // the winner's p / eps: pick the slot on every lane, then one shuffle from the owning lane
        // (explicit selp: written as an if-chain the compiler turns the slots into a local-memory array)
        float own_p = pk[0], own_e = ek[0];
#pragma unroll
        for (int e = 1; e < MAXE; ++e) {
          own_p = selp_f32((arg >> 5) == e, pk[e], own_p);
          own_e = selp_f32((arg >> 5) == e, ek[e], own_e);
        }
        const float psel = __shfl_sync(kFull, own_p, arg & 31);
        const float esel = __shfl_sync(kFull, own_e, arg & 31);
        const float g_a = ar.g_sample[r], g_u = ar.g_s_pre != nullptr ? ar.g_s_pre[r] : 0.f;
        // 1 - tanh(u)^2 without the fp32 cancellation of the literal form: sech^2(u) = 4 e^{-2|u|} / (1 + e^{-2|u|})^2
        float t, omt2, coef;
        if (FAST) {
          const float e2m = ex2f(-2.f * kLog2e * fabsf(psel)), q1 = rcpf(1.f + e2m);
          omt2 = 4.f * e2m * q1 * q1;
          t = copysignf(1.f - 2.f * e2m * q1, psel);  // tanh|u| = (1 - e^{-2|u|}) / (1 + e^{-2|u|})
          coef = fmaf(g_u, rcpf(fmaxf(1e-6f, omt2)), g_a);
        } else {
          const float e2m = expf(-2.f * fabsf(psel));
          omt2 = 4.f * e2m / ((1.f + e2m) * (1.f + e2m));
          t = tanhf(psel);
          coef = g_a + g_u / fmaxf(1e-6f, omt2);
        }
        // w = softmax(y); D_k = (tanh p_k - t) (g_a + g_u / max(1e-6, 1 - t^2)); dlogits = w (D - sum w D)
        float wexp[MAXE], D[MAXE];
        float s = 0.f, wd = 0.f;
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
          wexp[e] = 0.f;
          D[e] = 0.f;
          if (lane + 32 * e < P) {
            wexp[e] = FAST ? ex2f(y[e] - m) : expf(y[e] - m);
            D[e] = ((FAST ? fast_tanh(pk[e]) : tanhf(pk[e])) - t) * coef;
            s += wexp[e];
            wd = fmaf(wexp[e], D[e], wd);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s += __shfl_xor_sync(kFull, s, o);
          wd += __shfl_xor_sync(kFull, wd, o);
        }
        const float inv_s = FAST ? rcpf(s) : 1.f / s;
        wd *= inv_s;
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
          const int k = lane + 32 * e;
          if (k < P) ar.dlogits[r * P + k] = wexp[e] * inv_s * (D[e] - wd);
        }
        if (lane == (arg & 31)) {  // this warp is the only writer of row a of the table: plain read-modify-write
          const float gp = omt2 * g_a + g_u;                               // dL/dp_{k*}
          acc_s[a * P + arg] += gp;
          acc_s[AP + a * P + arg] += gp * (sd_s[a * P + arg] * esel);      // d p_{k*} / d logstd_{k*} = scale * eps
        }
      }
      if (BWD) {
        a += nw;
      } else if (++a == A) {
        a = 0;
      }
    }
    // ---- lane-parallel tail: lane j finishes row r0 + j ----
    if (!BWD && lane < nrow) {
      const long long r = r0 + lane;
      int al = a0 + lane;
      al -= (al / A) * A;
      {
        float eps;
        if (FAST) {
          // the forward only needs the winner's location draw: regenerate just that block
          const int es = my_arg >> 5;
          float n4[4];
          normal4(rng(rng_off, kNormalStream | (uint64_t)(r * P + (my_arg & 31) + 32 * (es & ~3))), n4);
          eps = (es & 2) ? ((es & 1) ? n4[3] : n4[2]) : ((es & 1) ? n4[1] : n4[0]);
        } else {
          eps = ar.ext_normal[r * P + my_arg];
        }
        const float lsd = ar.logstd[al * P + my_arg], lc = ar.loc[al * P + my_arg];
        const float psel = FAST ? fmaf(eps, ex2f(lsd * kLog2e), lc) : __fadd_rn(__fmul_rn(eps, expf(lsd)), lc);
        ar.sample[r] = tanhf(psel);
        ar.s_pre[r] = psel;
        ar.idx[r] = my_arg;
      }
    }
  }
  if (BWD) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * AP; i += blockDim.x) part[(size_t)blockIdx.x * 2 * AP + i] = acc_s[i];
  }
}

// dloc / dlogstd += the per-CTA tables, summed in CTA order (the caller pre-zeroes or pre-loads the outputs)
__global__ void rsample_bwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ dloc, float* __restrict__ dlogstd,
                                            int nparts, int AP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * AP) return;
  float t = 0.f;
  for (int c = 0; c < nparts; ++c) t += part[(size_t)c * 2 * AP + i];
  if (i < AP) dloc[i] += t;
  else dlogstd[i - AP] += t;
}

constexpr int kRsBwdMaxCtas = 148 * 4;
static int rs_bwd_warps(int A) { return A % 12 == 0 ? 12 : (A % 9 == 0 ? 9 : 8); }  // rows of a state split evenly

template <bool BWD>
static int rs_launch(const pfpn_rsample_args& a, size_t smem, float* part, cudaStream_t st) {
  const bool fast = a.ext_uniform == nullptr;
  const long long rows = (long long)a.B * a.A;
  // chunk = rows one warp walks before its lane-parallel tail: 32 at scale, fewer when the batch is
  // too small to fill 148 SMs x 4 CTAs x 8 warps with full chunks (acting-path latency)
  long long chunk = rows / (148LL * 4 * kRsWarps);
  chunk = chunk < 1 ? 1 : (chunk > 32 ? 32 : chunk);
  const long long nchunks = (rows + chunk - 1) / chunk;
  long long grid = (nchunks + kRsWarps - 1) / kRsWarps;
  if (grid > 148LL * 8) grid = 148LL * 8;
  const int threads = BWD ? rs_bwd_warps(a.A) * 32 : kRsWarps * 32;
  const int maxe = a.P <= 64 ? 2 : (a.P <= 128 ? 4 : 8);
#define PFPN_RS(F, E)                                                                                                   \
  do {                                                                                                                  \
    if (BWD) {                                                                                                          \
      PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)rsample_kernel<BWD, F, E>,                                         \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
      int occ = 0;                                                                                                      \
      PFPN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)rsample_kernel<BWD, F, E>, threads, \
                                                                 smem));                                                \
      grid = 148LL * (occ < 1 ? 1 : (occ > 4 ? 4 : occ));                                                               \
      if (grid > a.B) grid = a.B;                                                                                       \
    }                                                                                                                   \
    rsample_kernel<BWD, F, E><<<(int)grid, threads, smem, st>>>(a, (int)chunk, part);                                   \
  } while (0)
  if (fast) {
    if (maxe == 2) PFPN_RS(true, 2);
    else if (maxe == 4) PFPN_RS(true, 4);
    else PFPN_RS(true, 8);
  } else {
    if (maxe == 2) PFPN_RS(false, 2);
    else if (maxe == 4) PFPN_RS(false, 4);
    else PFPN_RS(false, 8);
  }
#undef PFPN_RS
  PFPN_CUDA_OK(cudaGetLastError());
  if (BWD) {
    const int AP = a.A * a.P;
    rsample_bwd_finalize_kernel<<<(2 * AP + 255) / 256, 256, 0, st>>>(part, a.dloc, a.dlogstd, (int)grid, AP);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  return PFPN_OK;
}

static int rsample_common(const pfpn_rsample_args& a) {
  if (a.B < 0 || a.A <= 0 || a.P <= 0) return PFPN_ERR_ARG;
  if (a.P > 256) return PFPN_ERR_UNSUPPORTED;
  if (!a.logits || !a.loc || !a.logstd) return PFPN_ERR_ARG;
  if ((a.ext_uniform == nullptr) != (a.ext_normal == nullptr)) return PFPN_ERR_ARG;
  return PFPN_OK;
}

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_head_rsample_fwd(const pfpn_rsample_args* args, pfpn_stream_t stream_) {
  if (!args) return PFPN_ERR_ARG;
  const pfpn_rsample_args& a = *args;
  int rc = rsample_common(a);
  if (rc != PFPN_OK) return rc;
  if (a.B == 0) return PFPN_OK;
  if (!a.sample || !a.s_pre || !a.idx) return PFPN_ERR_ARG;
  return rs_launch<false>(a, 0, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int pfpn_rsample_bwd_workspace_bytes(int32_t B, int32_t A, int32_t P, size_t* bytes) {
  if (!bytes || B < 0 || A <= 0 || P <= 0) return PFPN_ERR_ARG;
  *bytes = (size_t)kRsBwdMaxCtas * 2 * A * P * sizeof(float) + 16;
  return PFPN_OK;
}

extern "C" int pfpn_head_rsample_bwd(const pfpn_rsample_args* args, void* workspace, size_t workspace_bytes,
                                     pfpn_stream_t stream_) {
  if (!args) return PFPN_ERR_ARG;
  const pfpn_rsample_args& a = *args;
  int rc = rsample_common(a);
  if (rc != PFPN_OK) return rc;
  if (a.B == 0) return PFPN_OK;
  if (!a.g_sample || !a.dlogits || !a.dloc || !a.dlogstd) return PFPN_ERR_ARG;
  const size_t smem = 4 * (size_t)a.A * a.P * sizeof(float);  // acc[2], loc, sd
  if (smem > 200 * 1024) return PFPN_ERR_UNSUPPORTED;
  size_t need;
  pfpn_rsample_bwd_workspace_bytes(a.B, a.A, a.P, &need);
  if (!workspace || workspace_bytes < need) return PFPN_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return PFPN_ERR_ALIGN;
  return rs_launch<true>(a, smem, reinterpret_cast<float*>(workspace), reinterpret_cast<cudaStream_t>(stream_));
}
