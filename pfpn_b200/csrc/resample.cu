// K5 -- dead-particle resampling, and K4 -- running activity statistics.
//
// Replaces the ~120-node `cond/*` sub-graph that
// /root/reference/networks/actor_critic/a2c.py:385-474 builds (Where, Multinomial,
// GatherNd, UniqueWithCounts, map/while, ScatterNdUpdate ...) with three launches:
//   1. resample_plan_kernel  (one CTA): stable row-major compaction of the dead particles,
//      fp64-CDF categorical draw of the candidate table (TF-1.14 CPU Multinomial semantics),
//      per-dead-particle gathers, noise, logit split  b -= log(count + 1 - delta), and the
//      [A,P]-sized scatters;
//   2. resample_gather_w / 3. resample_scatter_w: fc_policy weight columns W[:, col_m] =
//      W_pre[:, tcol_m] through a snapshot, so a dead particle that is itself a source is
//      copied from its PRE-update value as the reference's gather-before-scatter does.
// Integer results (M, (a_m, j_m), cand, src, col, tcol, uniq/idx/count/delta) are bit-exact
// against oracle/resample.py; M is only known on the device, launches 2/3 size for A*P.
#include "common.cuh"

namespace pfpn {

struct ResampleWs {
  int* M;          // [1]
  int* nuniq;      // [1]
  int* invalid;    // [AP][2]
  int* cand;       // [A][P]
  int* src;        // [AP]
  int* col;        // [AP]
  int* tcol;       // [AP]
  int* count_by;   // [AP]  #m with tcol_m == column
  int* first_m;    // [AP]  smallest m with tcol_m == column
  int* is_dead;    // [AP]
  int* uniq_rank;  // [AP]  rank of a column among first occurrences
  float* tloc;     // [AP]
  float* tlogstd;  // [AP]
  float* tb;       // [AP]
  float* logit;    // [AP]
  double* cdf;     // [AP]
  double* total;   // [A]
  float* tW;       // [H][AP]
};

__device__ __forceinline__ int block_exclusive_scan(int flag, int* warp_sums, int* total_out) {
  // returns exclusive prefix of `flag` over the CTA in thread order; *total_out = CTA total
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag != 0);
  const int pre = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_sums[warp] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < nw; ++w) {
    const int v = warp_sums[w];
    if (w < warp) base += v;
    tot += v;
  }
  __syncthreads();
  *total_out = tot;
  return base + pre;
}

__global__ void __launch_bounds__(1024) resample_plan_kernel(const pfpn_resample_args ar, const ResampleWs ws) {
  __shared__ int warp_sums[32];
  const int A = ar.A, P = ar.P, AP = A * P;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float thr = ar.threshold > 0.f ? ar.threshold : 0.05f / (float)P;  // a2c.py:391
  const bool tanh_flag = (ar.flags & PFPN_RESAMPLE_FLAG_TANH) != 0;
  const int K = ar.resample < 0 ? P : min(P, ar.resample);
  const Philox rng(ar.seed);

  // ---- 1. avg = sum_active / rowsum (sequential fp32 row sum), logits = log(avg) ------------
  for (int a = tid; a < A; a += nthr) {
    float rs = 0.f;
    for (int k = 0; k < P; ++k) rs = __fadd_rn(rs, ar.sum_active[a * P + k]);
    double run = 0.0;
    float mx = -3.402823466e38f;
    for (int k = 0; k < P; ++k) {
      const float avg = __fdiv_rn(ar.sum_active[a * P + k], rs);
      const float lg = (float)log((double)avg);  // fp64 log rounded to fp32 (see oracle/resample.py)
      ws.logit[a * P + k] = ar.resample < 0 ? lg : avg;
      if (isfinite(lg)) mx = fmaxf(mx, lg);
    }
    if (ar.resample < 0) {  // TF CPU Multinomial: fp64 running CDF over the finite logits
      for (int k = 0; k < P; ++k) {
        const float lg = ws.logit[a * P + k];
        if (isfinite(lg)) run += exp((double)lg - (double)mx);
        ws.cdf[a * P + k] = run;
      }
      ws.total[a] = run;
    }
  }
  for (int i = tid; i < AP; i += nthr) {
    ws.count_by[i] = 0;
    ws.first_m[i] = 0x7fffffff;
    ws.is_dead[i] = 0;
  }
  __syncthreads();

  // ---- 2. candidate table ------------------------------------------------------------------
  if (ar.resample < 0) {
    for (int i = tid; i < A * P; i += nthr) {  // (a, s): s-th draw of row a
      const int a = i / P;
      double u;
      if (ar.ext_cat_u != nullptr) {
        u = ar.ext_cat_u[i];
      } else {
        const uint4 r = rng(ar.offset, (uint64_t)i);
        u = u64_to_unit_double(r.x, r.y);
      }
      const double to_find = u * ws.total[a];
      const double* cdf = ws.cdf + a * P;
      int lo = 0, hi = P;  // upper_bound: first index with cdf[idx] > to_find
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] > to_find) hi = mid;
        else lo = mid + 1;
      }
      ws.cand[i] = min(lo, P - 1);  // (all-zero statistics: TF would emit the out-of-range class P)
    }
  } else {
    for (int i = tid; i < AP; i += nthr) {  // descending rank, ties -> lower index
      const int a = i / P, k = i - a * P;
      const float v = ws.logit[i];
      int rank = 0;
      for (int j = 0; j < P; ++j) {
        const float w = ws.logit[a * P + j];
        rank += (w > v) || (w == v && j < k);
      }
      if (rank < K) ws.cand[a * K + rank] = k;
    }
  }

  // ---- 3. stable row-major compaction of dead particles (tf.where order) --------------------
  int M = 0;
  for (int base = 0; base < AP; base += nthr) {
    const int i = base + tid;
    const int flag = (i < AP) && (ar.max_active[i] < thr);
    int tot;
    const int pos = block_exclusive_scan(flag, warp_sums, &tot);
    if (flag) {
      const int m = M + pos;
      ws.invalid[2 * m] = i / P;
      ws.invalid[2 * m + 1] = i % P;
      ws.is_dead[i] = 1;
    }
    M += tot;
  }
  if (tid == 0) *ws.M = M;
  __syncthreads();

  // ---- 4. per dead particle: source, gathers of PRE-update values, noise ---------------------
  for (int m = tid; m < M; m += nthr) {
    const int a = ws.invalid[2 * m], j = ws.invalid[2 * m + 1];
    int ch;
    if (ar.resample < 0) {
      ch = j;  // a2c.py:403
    } else if (ar.ext_choice != nullptr) {
      ch = ar.ext_choice[m];
    } else {
      const uint4 r = rng(ar.offset + 1, (uint64_t)m);
      ch = (int)(r.x % (uint32_t)K);
    }
    const int s = ws.cand[a * K + ch];
    const int col = a * P + j, tcol = a * P + s;
    ws.src[m] = s;
    ws.col[m] = col;
    ws.tcol[m] = tcol;
    atomicAdd(&ws.count_by[tcol], 1);
    atomicMin(&ws.first_m[tcol], m);
    float tloc = ar.loc[tcol];
    const float tls = ar.logstd[tcol];
    const float tstd = expf(tls);
    float u;
    if (ar.ext_noise_u != nullptr) {
      u = ar.ext_noise_u[m];
    } else {
      const uint4 r = rng(ar.offset + 2, (uint64_t)m);
      u = 2.f * u32_to_unit_open(r.x) - 1.f;
    }
    float noise = __fmul_rn(tstd, u);
    noise = __fadd_rn(noise, noise < 0.f ? -1e-4f : 1e-4f);  // a2c.py:442-444
    tloc = __fadd_rn(tloc, noise);
    if (tanh_flag) {  // a2c.py:448-450
      const float eps = 1e-6f;
      tloc = atanhf(fminf(fmaxf(tloc, eps - 1.f), 1.f - eps));
    }
    ws.tloc[m] = tloc;
    ws.tlogstd[m] = fminf(fmaxf(tls, -20.f), 2.f);  // a2c.py:451
    ws.tb[m] = ar.bias[tcol];
  }
  __syncthreads();

  // ---- 5. unique_with_counts bookkeeping (first-occurrence order) + logit split -------------
  int nuniq = 0;
  for (int base = 0; base < M; base += nthr) {
    const int m = base + tid;
    const int flag = (m < M) && (ws.first_m[ws.tcol[m]] == m);
    int tot;
    const int pos = block_exclusive_scan(flag, warp_sums, &tot);
    if (flag) {
      const int u = nuniq + pos;
      const int tc = ws.tcol[m];
      ws.uniq_rank[tc] = u;
      if (ar.out_uniq) ar.out_uniq[u] = tc;
      if (ar.out_count) ar.out_count[u] = ws.count_by[tc];
      if (ar.out_delta) ar.out_delta[u] = ws.is_dead[tc];
    }
    nuniq += tot;
  }
  if (tid == 0) *ws.nuniq = nuniq;
  __syncthreads();
  for (int m = tid; m < M; m += nthr) {
    const int tc = ws.tcol[m];
    // a2c.py:458: b -= log(count + 1 - delta); delta = 1 iff the source column is itself dead
    const float denom = __fsub_rn(__fadd_rn((float)ws.count_by[tc], 1.f), (float)ws.is_dead[tc]);
    ws.tb[m] = __fsub_rn(ws.tb[m], logf(denom));
    if (ar.out_idx) ar.out_idx[m] = ws.uniq_rank[tc];
    if (ar.out_src) ar.out_src[m] = ws.src[m];
    if (ar.out_col) ar.out_col[m] = ws.col[m];
    if (ar.out_tcol) ar.out_tcol[m] = tc;
    if (ar.out_invalid) {
      ar.out_invalid[2 * m] = ws.invalid[2 * m];
      ar.out_invalid[2 * m + 1] = ws.invalid[2 * m + 1];
    }
  }
  if (ar.out_cand) {
    for (int i = tid; i < A * K; i += nthr) ar.out_cand[i] = ws.cand[i];
  }
  if (tid == 0) {
    if (ar.out_M) *ar.out_M = M;
    if (ar.out_nuniq) *ar.out_nuniq = nuniq;
  }
  __syncthreads();

  // ---- 6. [A,P]-sized scatters (a2c.py:460-466), then zero the statistics (a2c.py:372-378) ---
  for (int m = tid; m < M; m += nthr) {
    ar.loc[ws.col[m]] = ws.tloc[m];
    ar.logstd[ws.col[m]] = ws.tlogstd[m];
    ar.bias[ws.tcol[m]] = ws.tb[m];  // duplicates carry equal values
  }
  __syncthreads();
  for (int m = tid; m < M; m += nthr) ar.bias[ws.col[m]] = ws.tb[m];  // control-dependent second scatter
  for (int i = tid; i < AP; i += nthr) {
    ar.max_active[i] = 0.f;
    ar.sum_active[i] = 0.f;
  }
}

// tW[h][m] = W[h][tcol_m]   (snapshot of the source columns)
__global__ void resample_gather_w(const float* __restrict__ W, float* __restrict__ tW, const int* __restrict__ Mp,
                                  const int* __restrict__ tcol, int H, int AP) {
  const int M = *Mp;
  const int h = blockIdx.y;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x)
    tW[(size_t)h * AP + m] = W[(size_t)h * AP + tcol[m]];
}
// W[h][col_m] = tW[h][m]
__global__ void resample_scatter_w(float* __restrict__ W, const float* __restrict__ tW, const int* __restrict__ Mp,
                                   const int* __restrict__ col, int H, int AP) {
  const int M = *Mp;
  const int h = blockIdx.y;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x)
    W[(size_t)h * AP + col[m]] = tW[(size_t)h * AP + m];
}

// ---------------------------------------------------------------------------------------------
// K4: running activity statistics (a2c.py:346-365): max_active = max(max_active, max_b softmax(logits)),
// sum_active += sum_b softmax(logits).  Grid (batch chunks, A): every warp of CTA (c, a) walks rows (b, a) of its
// chunk with the row's particles in registers, so the per-(a,k) running max / sum live in registers -- no atomics,
// and the result is deterministic: warps are combined in order inside the CTA, chunks in order by the second kernel.
// Optionally writes the probabilities.
// ---------------------------------------------------------------------------------------------
constexpr int kStatsWarps = 8;
template <int MAXE>
__global__ void __launch_bounds__(kStatsWarps * 32) stats_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                                                 float* __restrict__ part, int B, int A, int P, int b_per_chunk) {
  __shared__ float sh[2][kStatsWarps][32 * MAXE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = blockIdx.y;
  const int b_lo = blockIdx.x * b_per_chunk, b_hi = min(B, b_lo + b_per_chunk);
  float vmax[MAXE], vsum[MAXE];
#pragma unroll
  for (int e = 0; e < MAXE; ++e) vmax[e] = vsum[e] = 0.f;
  constexpr int U = 4;  // rows in flight per warp: the rows of one (chunk, a) are A*P floats apart, latency-bound otherwise
  for (int b0 = b_lo + warp; b0 < b_hi; b0 += kStatsWarps * U) {
    float ex[U][MAXE];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int b = b0 + u * kStatsWarps;
      const float* x = logits + ((size_t)b * A + a) * P;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int k = lane + 32 * e;
        ex[u][e] = (b < b_hi && k < P) ? x[k] : -3.402823466e38f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int b = b0 + u * kStatsWarps;
      if (b >= b_hi) break;  // warp-uniform
      const size_t r = (size_t)b * A + a;
      float m = -3.402823466e38f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) m = fmaxf(m, ex[u][e]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        ex[u][e] = (lane + 32 * e < P) ? expf(ex[u][e] - m) : 0.f;
        s += ex[u][e];
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float inv = 1.f / s;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int k = lane + 32 * e;
        const float p = ex[u][e] * inv;
        if (k < P) {
          if (probs != nullptr) probs[r * P + k] = p;
          vmax[e] = fmaxf(vmax[e], p);
          vsum[e] += p;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < MAXE; ++e) {
    sh[0][warp][lane + 32 * e] = vmax[e];
    sh[1][warp][lane + 32 * e] = vsum[e];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < P; k += blockDim.x) {
    float mx = 0.f, sm = 0.f;
#pragma unroll
    for (int w = 0; w < kStatsWarps; ++w) {
      mx = fmaxf(mx, sh[0][w][k]);
      sm += sh[1][w][k];
    }
    float* dst = part + ((size_t)blockIdx.x * A + a) * P * 2;
    dst[k] = mx;
    dst[P + k] = sm;
  }
}
__global__ void stats_finalize_kernel(const float* __restrict__ part, float* __restrict__ max_active,
                                      float* __restrict__ sum_active, int nchunks, int A, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A * P) return;
  const int a = i / P, k = i - a * P;
  float mx = 0.f, sm = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const float* src = part + ((size_t)c * A + a) * P * 2;
    mx = fmaxf(mx, src[k]);
    sm += src[P + k];
  }
  max_active[i] = fmaxf(max_active[i], mx);
  sum_active[i] += sm;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t carve(ResampleWs* ws, unsigned char* base, int A, int P, int H) {
  const size_t AP = (size_t)A * P;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off = align_up(off + bytes, 16);
    return p;
  };
  ws->cdf = reinterpret_cast<double*>(take(AP * 8));
  ws->total = reinterpret_cast<double*>(take((size_t)A * 8));
  ws->M = reinterpret_cast<int*>(take(16));
  ws->nuniq = reinterpret_cast<int*>(take(16));
  ws->invalid = reinterpret_cast<int*>(take(AP * 8));
  ws->cand = reinterpret_cast<int*>(take(AP * 4));
  ws->src = reinterpret_cast<int*>(take(AP * 4));
  ws->col = reinterpret_cast<int*>(take(AP * 4));
  ws->tcol = reinterpret_cast<int*>(take(AP * 4));
  ws->count_by = reinterpret_cast<int*>(take(AP * 4));
  ws->first_m = reinterpret_cast<int*>(take(AP * 4));
  ws->is_dead = reinterpret_cast<int*>(take(AP * 4));
  ws->uniq_rank = reinterpret_cast<int*>(take(AP * 4));
  ws->tloc = reinterpret_cast<float*>(take(AP * 4));
  ws->tlogstd = reinterpret_cast<float*>(take(AP * 4));
  ws->tb = reinterpret_cast<float*>(take(AP * 4));
  ws->logit = reinterpret_cast<float*>(take(AP * 4));
  ws->tW = reinterpret_cast<float*>(take((size_t)H * AP * 4));
  return off;
}

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_resample_workspace_bytes(int32_t A, int32_t P, int32_t H, size_t* bytes) {
  if (!bytes || A <= 0 || P <= 0 || H < 0) return PFPN_ERR_ARG;
  ResampleWs ws;
  *bytes = carve(&ws, nullptr, A, P, H);
  return PFPN_OK;
}

extern "C" int pfpn_resample(const pfpn_resample_args* args, void* workspace, size_t workspace_bytes,
                             pfpn_stream_t stream_) {
  if (!args) return PFPN_ERR_ARG;
  const pfpn_resample_args& a = *args;
  if (a.A <= 0 || a.P <= 0 || a.H < 0) return PFPN_ERR_ARG;
  if (!a.max_active || !a.sum_active || !a.loc || !a.logstd || !a.bias || (a.H > 0 && !a.weight)) return PFPN_ERR_ARG;
  if (a.resample == 0 || a.resample < -1) return PFPN_ERR_ARG;
  if (a.resample > a.P) return PFPN_ERR_ARG;  // reference asserts n >= resample (a2c.py:390)
  ResampleWs ws;
  const size_t need = carve(&ws, nullptr, a.A, a.P, a.H);
  if (!workspace || workspace_bytes < need) return PFPN_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return PFPN_ERR_ALIGN;
  carve(&ws, reinterpret_cast<unsigned char*>(workspace), a.A, a.P, a.H);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  resample_plan_kernel<<<1, 1024, 0, stream>>>(a, ws);
  PFPN_CUDA_OK(cudaGetLastError());
  if (a.H > 0) {
    const int AP = a.A * a.P;
    dim3 grid((AP + 255) / 256, a.H);
    resample_gather_w<<<grid, 256, 0, stream>>>(a.weight, ws.tW, ws.M, ws.tcol, a.H, AP);
    PFPN_CUDA_OK(cudaGetLastError());
    resample_scatter_w<<<grid, 256, 0, stream>>>(a.weight, ws.tW, ws.M, ws.col, a.H, AP);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  return PFPN_OK;
}

static int stats_chunks(int B, int A) {
  int nb = (148 * 4 + A - 1) / A;  // ~4 CTAs per SM overall
  if (nb > B) nb = B;
  return nb < 1 ? 1 : nb;
}

extern "C" int pfpn_stats_workspace_bytes(int32_t B, int32_t A, int32_t P, size_t* bytes) {
  if (!bytes || B < 0 || A <= 0 || P <= 0) return PFPN_ERR_ARG;
  *bytes = (size_t)stats_chunks(B, A) * A * P * 2 * sizeof(float) + 16;
  return PFPN_OK;
}

extern "C" int pfpn_stats_update(const float* logits, float* probs, float* max_active, float* sum_active, int32_t B,
                                 int32_t A, int32_t P, void* workspace, size_t workspace_bytes, pfpn_stream_t stream_) {
  if (!logits || !max_active || !sum_active || B < 0 || A <= 0 || P <= 0) return PFPN_ERR_ARG;
  if (B == 0) return PFPN_OK;
  if (P > 256) return PFPN_ERR_UNSUPPORTED;
  size_t need;
  pfpn_stats_workspace_bytes(B, A, P, &need);
  if (!workspace || workspace_bytes < need) return PFPN_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return PFPN_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  float* part = reinterpret_cast<float*>(workspace);
  const int nb = stats_chunks(B, A);
  const int bpc = (B + nb - 1) / nb;
  const int nchunks = (B + bpc - 1) / bpc;
  dim3 grid(nchunks, A);
  if (P <= 64) stats_kernel<2><<<grid, kStatsWarps * 32, 0, st>>>(logits, probs, part, B, A, P, bpc);
  else if (P <= 128) stats_kernel<4><<<grid, kStatsWarps * 32, 0, st>>>(logits, probs, part, B, A, P, bpc);
  else stats_kernel<8><<<grid, kStatsWarps * 32, 0, st>>>(logits, probs, part, B, A, P, bpc);
  PFPN_CUDA_OK(cudaGetLastError());
  stats_finalize_kernel<<<(A * P + 255) / 256, 256, 0, st>>>(part, max_active, sum_active, nchunks, A, P);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}
