// K6 (anchor path) -- fp32 FFMA GEMMs of the 1024-512 MLP trunk with fused epilogues.
//
// Replaces the MatMul / Add / Relu6 / Relu6Grad / MatMul_grad nodes that
// /root/reference/networks/ops.py:82-118 (`fc_layer`) and tf.gradients build for
// actor/fc{1,2}, fc_policy and critic/fc{1,2,3}.  This is the parity anchor: plain fp32 FMAs
// with fp32 accumulation, so the trunk stays inside the 1e-5 norm-wise tolerance by
// construction.  128x128x16 CTA tiles, 8x8 register micro-tiles, double-buffered shared memory,
// 128-bit global and shared accesses; split-K (blockIdx.z) with a deterministic second-stage
// reduction for the weight gradients, whose reduction axis is the batch.
//
//   C[M,N] = op(A) * op(B)   TA = false: A is [M,K] row-major      TA = true: A is [K,M] row-major
//                            TB = false: B is [K,N] row-major      TB = true: B is [N,K] row-major
// Epilogues: EPI_NONE, EPI_BIAS (+bias[n]), EPI_BIAS_RELU6, EPI_MASK6 (multiply by 1[0 < H < 6],
// the Relu6Grad of a forward activation H[M,N]).
// All leading dimensions and M/N/K extents of contiguous axes must be multiples of 4 (the state
// dimension 197 is padded to 200 by the normaliser kernel).
#include "common.cuh"

namespace pfpn {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_RELU6 = 2, EPI_MASK6 = 3 };

struct GemmP {
  const float* A;
  const float* B;
  float* C;
  const float* bias;  // [N]
  const float* Hm;    // [M, ldh] forward activation for EPI_MASK6
  int M, N, K, lda, ldb, ldc, ldh;
  int k_chunk;        // K range per blockIdx.z (split-K); C then points at [splits][M][ldc]
};

// loads a (rows x BK) operand tile whose K axis is contiguous in memory: element (r, k) = P[r*ld + k]
// and stores it transposed as S[k][r]
__device__ __forceinline__ void load_kcontig(float4 (&reg)[2], const float* __restrict__ P, int ld, int r0, int rmax,
                                             int k0, int kmax, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + (tid >> 2) + 64 * i;
    const int k = k0 + (tid & 3) * 4;
    reg[i] = (r < rmax && k < kmax) ? __ldg(reinterpret_cast<const float4*>(P + (size_t)r * ld + k)) : make_float4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void store_kcontig(float (*S)[BM + PAD], const float4 (&reg)[2], int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (tid >> 2) + 64 * i;
    const int k = (tid & 3) * 4;
    S[k + 0][r] = reg[i].x;
    S[k + 1][r] = reg[i].y;
    S[k + 2][r] = reg[i].z;
    S[k + 3][r] = reg[i].w;
  }
}
// operand tile whose M/N axis is contiguous: element (r, k) = P[k*ld + r]
__device__ __forceinline__ void load_rcontig(float4 (&reg)[2], const float* __restrict__ P, int ld, int r0, int rmax,
                                             int k0, int kmax, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = k0 + (tid >> 5) + 8 * i;
    const int r = r0 + (tid & 31) * 4;
    reg[i] = (r < rmax && k < kmax) ? __ldg(reinterpret_cast<const float4*>(P + (size_t)k * ld + r)) : make_float4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void store_rcontig(float (*S)[BM + PAD], const float4 (&reg)[2], int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = (tid >> 5) + 8 * i;
    const int r = (tid & 31) * 4;
    *reinterpret_cast<float4*>(&S[k][r]) = reg[i];
  }
}

template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(const GemmP p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * p.k_chunk;
  const int ke = min(p.K, kb + p.k_chunk);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
    if (TA) load_rcontig(ra, p.A, p.lda, m0, p.M, k0, ke, tid);
    else load_kcontig(ra, p.A, p.lda, m0, p.M, k0, ke, tid);
    if (TB) load_kcontig(rb, p.B, p.ldb, n0, p.N, k0, ke, tid);
    else load_rcontig(rb, p.B, p.ldb, n0, p.N, k0, ke, tid);
  };
  auto sstore = [&](int buf) {
    if (TA) store_rcontig(As[buf], ra, tid);
    else store_kcontig(As[buf], ra, tid);
    if (TB) store_kcontig(Bs[buf], rb, tid);
    else store_rcontig(Bs[buf], rb, tid);
  };
  gload(kb);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kb; k0 < ke; k0 += BK) {
    const bool more = k0 + BK < ke;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  float* C = p.C + (size_t)blockIdx.z * p.M * p.ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + (jh ? 64 + tx * 4 : tx * 4);
      if (n >= p.N) continue;  // N % 4 == 0: a float4 is either fully inside or fully outside
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU6) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      if (EPI == EPI_BIAS_RELU6) {
        v.x = fminf(fmaxf(v.x, 0.f), 6.f); v.y = fminf(fmaxf(v.y, 0.f), 6.f);
        v.z = fminf(fmaxf(v.z, 0.f), 6.f); v.w = fminf(fmaxf(v.w, 0.f), 6.f);
      }
      if (EPI == EPI_MASK6) {  // Relu6Grad: pass where 0 < h < 6
        const float4 h = __ldg(reinterpret_cast<const float4*>(p.Hm + (size_t)m * p.ldh + n));
        v.x = (h.x > 0.f && h.x < 6.f) ? v.x : 0.f; v.y = (h.y > 0.f && h.y < 6.f) ? v.y : 0.f;
        v.z = (h.z > 0.f && h.z < 6.f) ? v.z : 0.f; v.w = (h.w > 0.f && h.w < 6.f) ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(C + (size_t)m * p.ldc + n) = v;
    }
  }
}

// out[i] = sum_z part[z][i]   (deterministic split-K second stage), float4 wide
__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, size_t n4, int splits,
                                     size_t stride) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 s = make_float4(0, 0, 0, 0);
  for (int z = 0; z < splits; ++z) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)z * stride) + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = s;
}

// column sums (bias gradients): part[z][n] = sum over a row chunk of Y[m, n]; then reduced.
// With `wrow` != NULL the rows are weighted: part[z][n] = sum_r wrow[r] * Y[r, n]  (X^T dy for N_out = 1)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ Y, int M, int N, int ldy, int rows_per,
                                                     float* __restrict__ part, const float* __restrict__ wrow) {
  const int n = blockIdx.x * 64 + (threadIdx.x & 63);
  const int sub = threadIdx.x >> 6;  // 4 row phases
  const int r0 = blockIdx.y * rows_per;
  const int r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + sub; r < r1; r += 4) s += __ldg(&Y[(size_t)r * ldy + n]) * (wrow ? __ldg(&wrow[r]) : 1.f);
  __shared__ float sh[4][64];
  sh[sub][threadIdx.x & 63] = s;
  __syncthreads();
  if (sub == 0 && n < N) part[(size_t)blockIdx.y * N + n] = sh[0][threadIdx.x] + sh[1][threadIdx.x] + sh[2][threadIdx.x] + sh[3][threadIdx.x];
}
__global__ void colsum_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int N, int chunks) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int z = 0; z < chunks; ++z) s += part[(size_t)z * N + n];
  out[n] = s;
}

// critic head fc3 (N = 1): y[m] = dot(X[m, :], w) + b   and its backward pieces
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ X, int M, int K, int ldx,
                                                     const float* __restrict__ w, const float* __restrict__ b,
                                                     float* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  float s = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(X + (size_t)warp * ldx + k));
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w + k));
    s += x.x * ww.x + x.y * ww.y + x.z * ww.z + x.w * ww.w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[warp] = s + b[0];
}
// dX[m, k] = dy[m] * w[k] * 1[0 < H[m,k] < 6]
__global__ void outer_mask_kernel(const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ H,
                                  float* __restrict__ dX, int M, int K, int ld) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  const int k4 = K / 4;
  if (i >= (size_t)M * k4) return;
  const int m = (int)(i / k4), k = (int)(i % k4) * 4;
  const float g = __ldg(&dy[m]);
  const float4 ww = __ldg(reinterpret_cast<const float4*>(w + k));
  const float4 h = __ldg(reinterpret_cast<const float4*>(H + (size_t)m * ld + k));
  float4 v;
  v.x = (h.x > 0.f && h.x < 6.f) ? g * ww.x : 0.f; v.y = (h.y > 0.f && h.y < 6.f) ? g * ww.y : 0.f;
  v.z = (h.z > 0.f && h.z < 6.f) ? g * ww.z : 0.f; v.w = (h.w > 0.f && h.w < 6.f) ? g * ww.w : 0.f;
  *reinterpret_cast<float4*>(dX + (size_t)m * ld + k) = v;
}

template <bool TA, bool TB>
static int launch_gemm(const GemmP& p, int epi, int splits, cudaStream_t st) {
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  switch (epi) {
    case EPI_NONE: sgemm_kernel<TA, TB, EPI_NONE><<<grid, 256, 0, st>>>(p); break;
    case EPI_BIAS: sgemm_kernel<TA, TB, EPI_BIAS><<<grid, 256, 0, st>>>(p); break;
    case EPI_BIAS_RELU6: sgemm_kernel<TA, TB, EPI_BIAS_RELU6><<<grid, 256, 0, st>>>(p); break;
    case EPI_MASK6: sgemm_kernel<TA, TB, EPI_MASK6><<<grid, 256, 0, st>>>(p); break;
    default: return PFPN_ERR_ARG;
  }
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

static bool mult4(int x) { return (x & 3) == 0; }
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pfpn

using namespace pfpn;

// Y[M,N] = act(X[M,K] * W[K,N] + b)            (fc_layer forward, ops.py:82-118)
extern "C" int pfpn_mlp_linear_fwd(const float* X, int32_t ldx, const float* W, const float* b, float* Y, int32_t ldy,
                                   int32_t M, int32_t K, int32_t N, int32_t relu6, pfpn_stream_t stream_) {
  if (!X || !W || !b || !Y || M < 0 || K <= 0 || N <= 0) return PFPN_ERR_ARG;
  if (M == 0) return PFPN_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (N == 1) {  // critic/fc3
    if (!mult4(K) || !mult4(ldx) || !al16(X) || !al16(W)) return PFPN_ERR_ALIGN;
    if (relu6) return PFPN_ERR_UNSUPPORTED;
    rowdot_kernel<<<(M * 32 + 255) / 256, 256, 0, st>>>(X, M, K, ldx, W, b, Y);
    PFPN_CUDA_OK(cudaGetLastError());
    return PFPN_OK;
  }
  if (!mult4(K) || !mult4(N) || !mult4(ldx) || !mult4(ldy) || !al16(X) || !al16(W) || !al16(Y) || !al16(b)) return PFPN_ERR_ALIGN;
  GemmP p{X, W, Y, b, nullptr, M, N, K, ldx, N, ldy, 0, K};
  return launch_gemm<false, false>(p, relu6 ? EPI_BIAS_RELU6 : EPI_BIAS, 1, st);
}

// dX[M,K] = (dY[M,N] * W[K,N]^T) .* relu6'(Hin[M,K])   (Hin == NULL: no mask)
extern "C" int pfpn_mlp_linear_bwd_input(const float* dY, int32_t ldy, const float* W, const float* Hin, float* dX,
                                         int32_t ldx, int32_t M, int32_t K, int32_t N, pfpn_stream_t stream_) {
  if (!dY || !W || !dX || M < 0 || K <= 0 || N <= 0) return PFPN_ERR_ARG;
  if (M == 0) return PFPN_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (N == 1) {
    if (!Hin) return PFPN_ERR_UNSUPPORTED;
    if (!mult4(K) || !mult4(ldx) || !al16(W) || !al16(Hin) || !al16(dX)) return PFPN_ERR_ALIGN;
    const size_t n4 = (size_t)M * (K / 4);
    outer_mask_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(dY, W, Hin, dX, M, K, ldx);
    PFPN_CUDA_OK(cudaGetLastError());
    return PFPN_OK;
  }
  if (!mult4(K) || !mult4(N) || !mult4(ldx) || !mult4(ldy) || !al16(dY) || !al16(W) || !al16(dX)) return PFPN_ERR_ALIGN;
  // C[M,K] = A[M,N] * B^T with B = W stored [K,N] row-major  -> TB = true (B given as [N'=K rows, K'=N cols])
  GemmP p{dY, W, dX, nullptr, Hin, M, K, N, ldy, N, ldx, ldx, N};
  return launch_gemm<false, true>(p, Hin ? EPI_MASK6 : EPI_NONE, 1, st);
}

extern "C" int pfpn_mlp_wgrad_workspace_bytes(int32_t M, int32_t K, int32_t N, size_t* bytes) {
  if (!bytes || M < 0 || K <= 0 || N <= 0) return PFPN_ERR_ARG;
  *bytes = ((size_t)64 * K * N + (size_t)1024 * N) * sizeof(float) + 256;
  return PFPN_OK;
}

// dW[K,N] = X[M,K]^T * dY[M,N],  db[N] = sum_m dY[m, :]      (reduction over the batch, split-K)
extern "C" int pfpn_mlp_linear_bwd_weight(const float* X, int32_t ldx, const float* dY, int32_t ldy, float* dW, float* db,
                                          int32_t M, int32_t K, int32_t N, void* workspace, size_t workspace_bytes,
                                          pfpn_stream_t stream_) {
  if (!X || !dY || !dW || !db || M <= 0 || K <= 0 || N <= 0) return PFPN_ERR_ARG;
  if (!mult4(K) || !mult4(ldx) || !al16(X) || !al16(dW)) return PFPN_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  size_t need;
  pfpn_mlp_wgrad_workspace_bytes(M, K, N, &need);
  if (!workspace || workspace_bytes < need) return PFPN_ERR_WORKSPACE;
  float* ws = reinterpret_cast<float*>(workspace);
  // bias gradient
  int chunks = (M + 255) / 256;
  if (chunks > 1024) chunks = 1024;
  const int rows_per = (M + chunks - 1) / chunks;
  chunks = (M + rows_per - 1) / rows_per;
  float* cpart = ws + (size_t)64 * K * N;
  colsum_kernel<<<dim3((N + 63) / 64, chunks), 256, 0, st>>>(dY, M, N, ldy, rows_per, cpart, nullptr);
  PFPN_CUDA_OK(cudaGetLastError());
  colsum_reduce_kernel<<<(N + 255) / 256, 256, 0, st>>>(cpart, db, N, chunks);
  PFPN_CUDA_OK(cudaGetLastError());
  if (N == 1) {  // critic/fc3: dW[k] = sum_m X[m,k] * dy[m] = dy-weighted column sums of X
    if (ldy != 1) return PFPN_ERR_UNSUPPORTED;
    float* xpart = ws;  // [chunks][K]
    colsum_kernel<<<dim3((K + 63) / 64, chunks), 256, 0, st>>>(X, M, K, ldx, rows_per, xpart, dY);
    PFPN_CUDA_OK(cudaGetLastError());
    colsum_reduce_kernel<<<(K + 255) / 256, 256, 0, st>>>(xpart, dW, K, chunks);
    PFPN_CUDA_OK(cudaGetLastError());
    return PFPN_OK;
  }
  if (!mult4(N) || !mult4(ldy) || !al16(dY)) return PFPN_ERR_ALIGN;
  const int tiles = ((K + BM - 1) / BM) * ((N + BN - 1) / BN);
  int splits = (148 * 2 + tiles - 1) / tiles;
  if (splits > 64) splits = 64;
  int k_chunk = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (M + k_chunk - 1) / k_chunk;
  // C[K_in, N] = A^T * B with A = X stored [M, K_in] (TA), B = dY stored [M, N]; reduction length = M
  GemmP p{X, dY, splits == 1 ? dW : ws, nullptr, nullptr, K, N, M, ldx, ldy, N, 0, k_chunk};
  int rc = launch_gemm<true, false>(p, EPI_NONE, splits, st);
  if (rc != PFPN_OK) return rc;
  if (splits > 1) {
    const size_t n4 = (size_t)K * N / 4;
    splitk_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(ws, dW, n4, splits, (size_t)K * N);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  return PFPN_OK;
}

// db[N] = column sums of dY[M,N] (bias gradient alone; the tensor-core weight-gradient path uses it)
extern "C" int pfpn_bias_grad(const float* dY, int32_t ldy, float* db, int32_t M, int32_t N, void* workspace,
                              size_t workspace_bytes, pfpn_stream_t stream_) {
  if (!dY || !db || M <= 0 || N <= 0) return PFPN_ERR_ARG;
  if (!workspace || workspace_bytes < (size_t)1024 * N * sizeof(float)) return PFPN_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  int chunks = (M + 255) / 256;
  if (chunks > 1024) chunks = 1024;
  const int rows_per = (M + chunks - 1) / chunks;
  chunks = (M + rows_per - 1) / rows_per;
  float* cpart = reinterpret_cast<float*>(workspace);
  colsum_kernel<<<dim3((N + 63) / 64, chunks), 256, 0, st>>>(dY, M, N, ldy, rows_per, cpart, nullptr);
  PFPN_CUDA_OK(cudaGetLastError());
  colsum_reduce_kernel<<<(N + 255) / 256, 256, 0, st>>>(cpart, db, N, chunks);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}
