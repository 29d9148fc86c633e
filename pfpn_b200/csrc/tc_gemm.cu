// K6 (tensor-core path) -- 3xTF32 error-compensated GEMM on tcgen05 with TMA-fed operands and
// a TMEM accumulator, same epilogues as the FFMA anchor (csrc/sgemm.cu).
//
//   C[M,N] = A[M,K] * B[N,K]^T      fp32 operands, each either K-major (row-major [rows, K]) or
//                                   MN-major (row-major [K, rows]) -- template flags A_MN / B_MN.
// With both majors available no operand is ever transposed in HBM: forward X*W (A K-major, W as
// stored = MN-major B), input gradient dY*W^T (K-major, K-major), weight gradient X^T*dY (MN, MN).
//
// Plain TF32 (10-bit mantissa) cannot hold 1e-5 through three layers, so every fp32 operand is
// split on the fly into hi = tf32(x) and lo = x - hi (exact) and three MMAs are accumulated:
// hi*hi + hi*lo + lo*hi  (relative error ~2^-21 per product, fp32 accumulation in TMEM).
//
// CTA = 8 warps, one 128x128 output tile, K in blocks of 32 floats (= one 128-byte swizzle atom):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) of A and B into a TC_STAGES-deep ring
//   warp 1      single-thread tcgen05.mma issuer (kind::tf32, M=128, N=128, K=8), tcgen05.commit
//   warp 2      TMEM allocator (all 512 columns: two alternating main accumulators + one correction)
//   warps 4-7   splitter: thread m moves row m of the landed A tile into TENSOR MEMORY as (hi, lo) with
//               tcgen05.st -- the MMAs then take A from TMEM and only B from shared memory, which is
//               the scarce resource of this kernel -- and writes the B_lo tile elementwise in the
//               swizzled layout; afterwards the epilogue: tcgen05.ld -> bias / relu6 / Relu6Grad
//               mask -> global stores
// Pipelines: full[s] (TMA -> splitter), conv[s] (splitter -> MMA), empty[s] (MMA -> TMA),
// tmem_full (MMA -> epilogue).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace pfpn {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 4;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;                       // 16 KiB per operand tile
constexpr int TC_STAGE_BYTES = 3 * TC_TILE_BYTES;                     // A (raw), B (raw = hi), B_lo
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 + 256;  // + alignment slack + barriers
// TMEM columns: [0,128) [128,256) main accumulators (even / odd k-blocks), [256,384) correction
// accumulator, [384,512) the A operand: two slots of (hi: 32 columns, lo: 32 columns)
constexpr uint32_t TC_TMEM_A = 3 * TC_BN;
enum { TC_EPI_NONE = 0, TC_EPI_BIAS = 1, TC_EPI_BIAS_RELU6 = 2, TC_EPI_MASK6 = 3 };

struct TcParams {
  float* C;
  const float* bias;
  const float* Hm;
  int M, N, K, ldc, ldh, epi;
  int k_chunk;  // split-K: K range per blockIdx.z (multiple of TC_BK); C then is [splits][M][ldc]
  float* colsum;  // weight gradient only (B MN-major): [splits][N] column sums of the B operand (= bias gradient)
  uint32_t wait_ns;  // persistent kernel: back-off of the TMA and store warps' long waits
  int b_lo_tma;   // the B operand's low part (x - tf32(x)) exists in global memory (weights: split once per optimizer
                  // step) and arrives by TMA through mapBlo; the splitter then only converts the A operand
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// MN-major operand tile: four panels of [32 k-rows][32 floats along M/N] (what one TMA box of a row-major
// [K, rows] matrix lands as).  For 32-bit MN-major operands the only swizzled layout the tensor core takes
// is SWIZZLE_128B with 32-byte atomicity (layout type 1; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): the
// 32-byte chunk index is XORed with the row index mod 4, so a swizzle atom is 4 k-rows of 128 B.
// Canonical form ((4,8,m),(4,k)):((1,4,LBO),(32,SBO)) in floats: SBO = next 4 k-rows (512 B),
// LBO = next 32-float panel (4096 B).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand from tensor memory (lane = row m, 32-bit column = k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256, 1) tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                         const __grid_constant__ CUtensorMap mapB,
                                                         const __grid_constant__ CUtensorMap mapBlo, const TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gbase = smem_dyn + (base - raw);
  const uint32_t bar0 = base + TC_STAGES * TC_STAGE_BYTES;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto conv = [&](int s) { return bar0 + 8 * (TC_STAGES + s); };
  auto empty = [&](int s) { return bar0 + 8 * (2 * TC_STAGES + s); };
  const uint32_t tmem_full = bar0 + 8 * 3 * TC_STAGES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + TC_STAGES * TC_STAGE_BYTES + 8 * (3 * TC_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int k_begin = blockIdx.z * p.k_chunk;
  const int k_end = min(p.K, k_begin + p.k_chunk);
  const int nkb = (k_end - k_begin + TC_BK - 1) / TC_BK;  // (TMA zero-fills beyond K; chunks are TC_BK multiples)

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 128);
      mbar_init(empty(s), 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    if (p.b_lo_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBlo)) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        mbar_wait(empty(s), (uint32_t)(((kb / TC_STAGES) & 1) ^ 1));
        mbar_expect_tx(full(s), (p.b_lo_tma ? 3 : 2) * TC_TILE_BYTES);
        const uint32_t a_dst = base + s * TC_STAGE_BYTES, b_dst = a_dst + TC_TILE_BYTES;
        const int k0 = k_begin + kb * TC_BK;
        if (A_MN) {
#pragma unroll
          for (int q = 0; q < TC_BM / 32; ++q) tma_load_2d(a_dst + q * 4096, &mapA, m0 + 32 * q, k0, full(s));
        } else {
          tma_load_2d(a_dst, &mapA, k0, m0, full(s));
        }
        if (B_MN) {
#pragma unroll
          for (int q = 0; q < TC_BN / 32; ++q) tma_load_2d(b_dst + q * 4096, &mapB, n0 + 32 * q, k0, full(s));
        } else {
          tma_load_2d(b_dst, &mapB, k0, n0, full(s));
        }
        if (p.b_lo_tma) {
          const uint32_t l_dst = b_dst + TC_TILE_BYTES;
          if (B_MN) {
#pragma unroll
            for (int q = 0; q < TC_BN / 32; ++q) tma_load_2d(l_dst + q * 4096, &mapBlo, n0 + 32 * q, k0, full(s));
          } else {
            tma_load_2d(l_dst, &mapBlo, k0, n0, full(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B tf32, A from TMEM (always K-major), B major bit 16 (1 = MN-major), N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        mbar_wait(conv(s), (uint32_t)((kb / TC_STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_hi = base + s * TC_STAGE_BYTES + TC_TILE_BYTES, b_lo = b_hi + TC_TILE_BYTES;
        const uint32_t a_hi = tmem + TC_TMEM_A + (uint32_t)((kb & 1) * 64), a_lo = a_hi + 32;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          // one MMA eats 8 tf32 along K: 8 TMEM columns of A; for B 32 bytes inside the swizzle atom (K-major)
          // or two 4-row atoms (MN-major)
          const uint64_t db_hi = B_MN ? umma_desc_mn(b_hi + ks * 1024) : umma_desc(b_hi + ks * 32);
          const uint64_t db_lo = B_MN ? umma_desc_mn(b_lo + ks * 1024) : umma_desc(b_lo + ks * 32);
          // The tensor core truncates its fp32 accumulator on every MMA (measured: ~2^-25 relative per
          // accumulation, systematic).  The dominant hi*hi products therefore get their own accumulator
          // (K/8 accumulations) and the two small correction products a second one, whose truncation is
          // relative to its 2^-11 times smaller magnitude; the epilogue adds the two in fp32.
          // Even / odd k-blocks alternate between two main accumulators, halving the chain again.
          const uint32_t dmain = tmem + (uint32_t)((kb & 1) * TC_BN);
          umma_tf32_ts(dmain, a_hi + ks * 8, db_hi, idesc, (kb >= 2 || ks != 0) ? 1u : 0u);
          umma_tf32_ts(tmem + 2 * TC_BN, a_lo + ks * 8, db_hi, idesc, (kb | ks) != 0);
          umma_tf32_ts(tmem + 2 * TC_BN, a_hi + ks * 8, db_lo, idesc, 1u);
        }
        umma_commit(empty(s));  // implicit tcgen05.fence::before_thread_sync
      }
      umma_commit(tmem_full);
    }
  } else if (warp >= 4) {
    // -------- splitter ------------------------------------------------------------------------------
    // A: thread t owns row t of the tile (= TMEM lane t): x -> TMEM (hi = x as is: the tensor core ignores
    // the low 13 mantissa bits; lo = x - tf32(x), exact).  B: lo tile written elementwise in place-layout.
    const int t = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // bias gradient riding on the weight-gradient GEMM: the B operand (dY) passes through this thread's
    // registers anyway.  float4 number t + 128 j of an MN-major tile sits in panel j / 2 at a column
    // offset that depends only on t (the 32-byte-chunk swizzle uses (row & 3) = (t / 8) & 3), so four
    // float4 accumulators per thread cover its share; 16 threads share a column and are combined in a
    // fixed order after the main loop.
    const bool do_cs = B_MN && p.colsum != nullptr && blockIdx.y == 0;
    float4 cs[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) cs[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES;
      mbar_wait(full(s), (uint32_t)((kb / TC_STAGES) & 1));
      const unsigned char* sa = gbase + (size_t)s * TC_STAGE_BYTES;
      uint32_t xh[32], xl[32];
      if (A_MN) {
        // [4 panels][32 k-rows][32 floats along m], 32-byte chunks XORed with (k-row & 3)
        const unsigned char* pa = sa + (t >> 5) * 4096 + (t & 7) * 4;
        const int c = (t & 31) >> 3;
#pragma unroll
        for (int r = 0; r < 32; ++r) xh[r] = *reinterpret_cast<const uint32_t*>(pa + r * 128 + ((c ^ (r & 3)) << 5));
      } else {
        // [128 rows][32 floats along k], 16-byte chunks XORed with (row & 7)
        const unsigned char* pa = sa + t * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(pa + ((c ^ (t & 7)) << 4));
          xh[4 * c] = v.x; xh[4 * c + 1] = v.y; xh[4 * c + 2] = v.z; xh[4 * c + 3] = v.w;
        }
      }
#pragma unroll
      for (int r = 0; r < 32; ++r)
        xl[r] = __float_as_uint(__uint_as_float(xh[r]) - __uint_as_float(xh[r] & 0xffffe000u));
      // B_lo first: it only touches shared memory, so it overlaps the MMAs that still read the TMEM slot
      if (!p.b_lo_tma) {
        const float4* hi = reinterpret_cast<const float4*>(sa + TC_TILE_BYTES);
        float4* lo = reinterpret_cast<float4*>(const_cast<unsigned char*>(sa) + 2 * TC_TILE_BYTES);
#pragma unroll
        for (int j = 0; j < TC_TILE_BYTES / 16 / 128; ++j) {
          const int i = t + 128 * j;
          const float4 x = hi[i];
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          lo[i] = l;
          if (B_MN && do_cs) {
            cs[j >> 1].x += x.x; cs[j >> 1].y += x.y; cs[j >> 1].z += x.z; cs[j >> 1].w += x.w;
          }
        }
      }
      // the TMEM slot (kb & 1) was last read by the MMAs of k-block kb - 2: wait for their commit
      if (kb >= 2) mbar_wait(empty((kb - 2) % TC_STAGES), (uint32_t)((((kb - 2) / TC_STAGES)) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem + lane_base + TC_TMEM_A + (uint32_t)((kb & 1) * 64);
      tmem_st32(ta, xh);
      tmem_st32(ta + 32, xl);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_async_smem();  // generic-proxy writes of B_lo -> visible to the tensor core (async proxy)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(conv(s));
    }
    // -------- epilogue: TMEM -> registers -> global --------------------------------------------
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (B_MN && do_cs) {  // scratch behind the 64 KiB output staging area: [16 row groups][128 columns]
      float* scr = reinterpret_cast<float*>(gbase + 4 * TC_TILE_BYTES);
      const int noff = ((((t & 7) >> 1) ^ ((t >> 3) & 3)) << 3) + ((t & 1) << 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(scr + (t >> 3) * 128 + q * 32 + noff) = cs[q];
    }
    const int wq = warp & 3;              // TMEM lane quadrant this warp may read
#pragma unroll 1
    for (int c0 = 0; c0 < TC_BN; c0 += 32) {
      uint32_t r[32], q[32];
      const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
            "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
            "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
            "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
          : "r"(taddr + (uint32_t)(2 * TC_BN)));
      uint32_t o[32];
      if (nkb > 1) {  // the odd-k-block accumulator exists only when there are at least two k-blocks
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]),
              "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15]), "=r"(o[16]),
              "=r"(o[17]), "=r"(o[18]), "=r"(o[19]), "=r"(o[20]), "=r"(o[21]), "=r"(o[22]), "=r"(o[23]), "=r"(o[24]),
              "=r"(o[25]), "=r"(o[26]), "=r"(o[27]), "=r"(o[28]), "=r"(o[29]), "=r"(o[30]), "=r"(o[31])
            : "r"(taddr + (uint32_t)TC_BN));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = 0u;
      }
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // stage the summed row chunk in shared memory (the operand ring is idle by now): 128-byte rows,
      // 16-byte chunks XORed with (row & 7) -> conflict-free for this row-per-thread write and for the
      // row-per-warp read below
      unsigned char* srow = gbase + (size_t)(c0 >> 5) * TC_TILE_BYTES + (size_t)(wq * 32 + lane) * 128;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 v = make_float4((__uint_as_float(r[j]) + __uint_as_float(o[j])) + __uint_as_float(q[j]),
                                     (__uint_as_float(r[j + 1]) + __uint_as_float(o[j + 1])) + __uint_as_float(q[j + 1]),
                                     (__uint_as_float(r[j + 2]) + __uint_as_float(o[j + 2])) + __uint_as_float(q[j + 2]),
                                     (__uint_as_float(r[j + 3]) + __uint_as_float(o[j + 3])) + __uint_as_float(q[j + 3]));
        *reinterpret_cast<float4*>(srow + ((((j >> 2) ^ (lane & 7))) << 4)) = v;
      }
    }
  }
  // ---- coalesced second phase, all 8 warps (producer / MMA / allocator warps are idle by now): one warp
  // writes one full 512-byte output row per instruction; 8 rows per warp in flight hide the Hm load latency
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (B_MN && p.colsum != nullptr && blockIdx.y == 0 && threadIdx.x < TC_BN) {
    const float* scr = reinterpret_cast<const float*>(gbase + 4 * TC_TILE_BYTES);
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) sum += scr[g * 128 + threadIdx.x];
    if (n0 + (int)threadIdx.x < p.N) p.colsum[(size_t)blockIdx.z * p.N + n0 + threadIdx.x] = sum;
  }
  {
    const int n = n0 + lane * 4;
    const bool n_ok = n < p.N;  // N % 4 == 0
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_ok && (p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_RELU6)) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const unsigned char* sbuf = gbase + (size_t)(lane >> 3) * TC_TILE_BYTES;
    const int rows_here = min(TC_BM, p.M - m0);
    float* crow = p.C + ((size_t)blockIdx.z * p.M + m0) * p.ldc + n;
    if (n_ok) {
#pragma unroll 1
      for (int r0 = warp; r0 < rows_here; r0 += 64) {
        float4 h[8];
        if (p.epi == TC_EPI_MASK6) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int row = r0 + 8 * u;
            h[u] = row < rows_here ? __ldg(reinterpret_cast<const float4*>(p.Hm + (size_t)(m0 + row) * p.ldh + n))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int row = r0 + 8 * u;
          if (row >= rows_here) break;
          float4 v = *reinterpret_cast<const float4*>(sbuf + (size_t)row * 128 + ((((lane & 7) ^ (row & 7))) << 4));
          v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          if (p.epi == TC_EPI_BIAS_RELU6) {
            v.x = fminf(fmaxf(v.x, 0.f), 6.f); v.y = fminf(fmaxf(v.y, 0.f), 6.f);
            v.z = fminf(fmaxf(v.z, 0.f), 6.f); v.w = fminf(fmaxf(v.w, 0.f), 6.f);
          }
          if (p.epi == TC_EPI_MASK6) {
            v.x = (h[u].x > 0.f && h[u].x < 6.f) ? v.x : 0.f; v.y = (h[u].y > 0.f && h[u].y < 6.f) ? v.y : 0.f;
            v.z = (h[u].z > 0.f && h[u].z < 6.f) ? v.z : 0.f; v.w = (h[u].w > 0.f && h[u].w < 6.f) ? v.w : 0.f;
          }
          *reinterpret_cast<float4*>(crow + (size_t)row * p.ldc) = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent form (round 2): one CTA per SM walks the output tiles (tile = z * tiles_m * tiles_n + m * tiles_n + n).
// The operand ring, the TMEM allocation and the barriers live for the whole kernel, the TMA producer prefetches the next
// tile's k-blocks while the current tile drains, and the epilogue is split: warps 4-7 (the splitters) move the three
// accumulators from TMEM into a dedicated 64 KiB staging area and hand the accumulators back to the MMA warp at once;
// warps 8-11 then apply bias / relu6 / Relu6Grad mask and write the tile to global memory WHILE the next tile's main
// loop runs.  (A second accumulator set does not fit: 3 x 128 accumulator columns + 128 A-operand columns = 512.)
// Barriers: full / conv / empty per ring stage (phases follow a global k-block counter), acc_full (MMA -> splitters),
// acc_free (splitters -> MMA), stage_full (splitters -> store warps), stage_free (store warps -> splitters).
// ------------------------------------------------------------------------------------------------
constexpr int TCP_STAGES = 3;
constexpr int TCP_RING_BYTES = TCP_STAGES * TC_STAGE_BYTES;
constexpr int TCP_STAGING_BYTES = 4 * TC_TILE_BYTES;   // 128 x 128 fp32
constexpr int TCP_SCR_BYTES = 16 * 128 * 4;            // bias-gradient scratch
constexpr int TCP_SMEM_BYTES = TCP_RING_BYTES + TCP_STAGING_BYTES + TCP_SCR_BYTES + 1024 + 256;
constexpr int TCP_THREADS = 384;

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TCP_THREADS, 1) tc_gemm_persist_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                         const __grid_constant__ CUtensorMap mapB,
                                                                         const __grid_constant__ CUtensorMap mapBlo,
                                                                         const TcParams p, const int tiles_m, const int tiles_n,
                                                                         const int num_tiles) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - raw);
  unsigned char* staging = gbase + TCP_RING_BYTES;
  float* scr = reinterpret_cast<float*>(staging + TCP_STAGING_BYTES);
  const uint32_t bar0 = base + TCP_RING_BYTES + TCP_STAGING_BYTES + TCP_SCR_BYTES;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto conv = [&](int s) { return bar0 + 8 * (TCP_STAGES + s); };
  auto empty = [&](int s) { return bar0 + 8 * (2 * TCP_STAGES + s); };
  const uint32_t acc_full = bar0 + 8 * 3 * TCP_STAGES, acc_free = acc_full + 8, stage_full = acc_full + 16,
                 stage_free = acc_full + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + TCP_RING_BYTES + TCP_STAGING_BYTES + TCP_SCR_BYTES +
                                                                       8 * (3 * TCP_STAGES + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TCP_STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 128);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 128);
    mbar_init(stage_full, 128);
    mbar_init(stage_free, 128);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    if (p.b_lo_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBlo)) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int tiles_mn = tiles_m * tiles_n;
  auto decode = [&](int t, int& z, int& m0, int& n0, int& k_begin, int& nkb) {
    z = t / tiles_mn;
    const int rem = t - z * tiles_mn;
    const int mb = rem / tiles_n;
    m0 = mb * TC_BM;
    n0 = (rem - mb * tiles_n) * TC_BN;
    k_begin = z * p.k_chunk;
    const int k_end = min(p.K, k_begin + p.k_chunk);
    nkb = (k_end - k_begin + TC_BK - 1) / TC_BK;
  };

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int z, m0, n0, k_begin, nkb;
        decode(t, z, m0, n0, k_begin, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % TCP_STAGES;
          mbar_wait_sleep(empty(s), (uint32_t)(((g / TCP_STAGES) & 1) ^ 1), p.wait_ns);
          mbar_expect_tx(full(s), (p.b_lo_tma ? 3 : 2) * TC_TILE_BYTES);
          const uint32_t a_dst = base + s * TC_STAGE_BYTES, b_dst = a_dst + TC_TILE_BYTES, l_dst = b_dst + TC_TILE_BYTES;
          const int k0 = k_begin + kb * TC_BK;
          if (A_MN) {
#pragma unroll
            for (int q = 0; q < TC_BM / 32; ++q) tma_load_2d(a_dst + q * 4096, &mapA, m0 + 32 * q, k0, full(s));
          } else {
            tma_load_2d(a_dst, &mapA, k0, m0, full(s));
          }
          if (B_MN) {
#pragma unroll
            for (int q = 0; q < TC_BN / 32; ++q) tma_load_2d(b_dst + q * 4096, &mapB, n0 + 32 * q, k0, full(s));
          } else {
            tma_load_2d(b_dst, &mapB, k0, n0, full(s));
          }
          if (p.b_lo_tma) {
            if (B_MN) {
#pragma unroll
              for (int q = 0; q < TC_BN / 32; ++q) tma_load_2d(l_dst + q * 4096, &mapBlo, n0 + 32 * q, k0, full(s));
            } else {
              tma_load_2d(l_dst, &mapBlo, k0, n0, full(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int g = 0, ti = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++ti) {
        int z, m0, n0, k_begin, nkb;
        decode(t, z, m0, n0, k_begin, nkb);
        if (ti > 0) mbar_wait(acc_free, (uint32_t)((ti - 1) & 1));  // the previous tile's accumulators were read out
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % TCP_STAGES;
          mbar_wait(conv(s), (uint32_t)((g / TCP_STAGES) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b_hi = base + s * TC_STAGE_BYTES + TC_TILE_BYTES, b_lo = b_hi + TC_TILE_BYTES;
          const uint32_t a_hi = tmem + TC_TMEM_A + (uint32_t)((g & 1) * 64), a_lo = a_hi + 32;
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            const uint64_t db_hi = B_MN ? umma_desc_mn(b_hi + ks * 1024) : umma_desc(b_hi + ks * 32);
            const uint64_t db_lo = B_MN ? umma_desc_mn(b_lo + ks * 1024) : umma_desc(b_lo + ks * 32);
            const uint32_t dmain = tmem + (uint32_t)((kb & 1) * TC_BN);
            umma_tf32_ts(dmain, a_hi + ks * 8, db_hi, idesc, (kb >= 2 || ks != 0) ? 1u : 0u);
            umma_tf32_ts(tmem + 2 * TC_BN, a_lo + ks * 8, db_hi, idesc, (kb | ks) != 0);
            umma_tf32_ts(tmem + 2 * TC_BN, a_hi + ks * 8, db_lo, idesc, 1u);
          }
          umma_commit(empty(s));
        }
        umma_commit(acc_full);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int t128 = threadIdx.x - 128;
    const int wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    int g = 0, ti = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++ti) {
      int z, m0, n0, k_begin, nkb;
      decode(t, z, m0, n0, k_begin, nkb);
      const bool do_cs = B_MN && p.colsum != nullptr && m0 == 0;
      float4 cs[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) cs[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const int s = g % TCP_STAGES;
        mbar_wait(full(s), (uint32_t)((g / TCP_STAGES) & 1));
        const unsigned char* sa = gbase + (size_t)s * TC_STAGE_BYTES;
        uint32_t xh[32], xl[32];
        if (A_MN) {
          const unsigned char* pa = sa + (t128 >> 5) * 4096 + (t128 & 7) * 4;
          const int c = (t128 & 31) >> 3;
#pragma unroll
          for (int r = 0; r < 32; ++r) xh[r] = *reinterpret_cast<const uint32_t*>(pa + r * 128 + ((c ^ (r & 3)) << 5));
        } else {
          const unsigned char* pa = sa + t128 * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(pa + ((c ^ (t128 & 7)) << 4));
            xh[4 * c] = v.x; xh[4 * c + 1] = v.y; xh[4 * c + 2] = v.z; xh[4 * c + 3] = v.w;
          }
        }
#pragma unroll
        for (int r = 0; r < 32; ++r)
          xl[r] = __float_as_uint(__uint_as_float(xh[r]) - __uint_as_float(xh[r] & 0xffffe000u));
        if (!p.b_lo_tma) {
          const float4* hi = reinterpret_cast<const float4*>(sa + TC_TILE_BYTES);
          float4* lo = reinterpret_cast<float4*>(const_cast<unsigned char*>(sa) + 2 * TC_TILE_BYTES);
#pragma unroll
          for (int j = 0; j < TC_TILE_BYTES / 16 / 128; ++j) {
            const int i = t128 + 128 * j;
            const float4 x = hi[i];
            float4 l;
            l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            lo[i] = l;
            if (B_MN && do_cs) {
              cs[j >> 1].x += x.x; cs[j >> 1].y += x.y; cs[j >> 1].z += x.z; cs[j >> 1].w += x.w;
            }
          }
        }
        if (g >= 2) mbar_wait(empty((g - 2) % TCP_STAGES), (uint32_t)(((g - 2) / TCP_STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem + lane_base + TC_TMEM_A + (uint32_t)((g & 1) * 64);
        tmem_st32(ta, xh);
        tmem_st32(ta + 32, xl);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_async_smem();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(conv(s));
      }
      // ---- epilogue, first half: TMEM -> staging; then the accumulators are free again -------------------
      mbar_wait(acc_full, (uint32_t)(ti & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (ti > 0) mbar_wait(stage_free, (uint32_t)((ti - 1) & 1));  // the store warps are done with the previous tile
      if (B_MN && do_cs) {
        const int noff = ((((t128 & 7) >> 1) ^ ((t128 >> 3) & 3)) << 3) + ((t128 & 1) << 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(scr + (t128 >> 3) * 128 + q * 32 + noff) = cs[q];
      }
#pragma unroll 1
      for (int c0 = 0; c0 < TC_BN; c0 += 32) {
        uint32_t r[32], q[32], o[32];
        const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
              "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
              "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
              "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
            : "r"(taddr + (uint32_t)(2 * TC_BN)));
        if (nkb > 1) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]),
                "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15]), "=r"(o[16]),
                "=r"(o[17]), "=r"(o[18]), "=r"(o[19]), "=r"(o[20]), "=r"(o[21]), "=r"(o[22]), "=r"(o[23]), "=r"(o[24]),
                "=r"(o[25]), "=r"(o[26]), "=r"(o[27]), "=r"(o[28]), "=r"(o[29]), "=r"(o[30]), "=r"(o[31])
              : "r"(taddr + (uint32_t)TC_BN));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0u;
        }
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        unsigned char* srow = staging + (size_t)(c0 >> 5) * TC_TILE_BYTES + (size_t)(wq * 32 + lane) * 128;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 v = make_float4((__uint_as_float(r[j]) + __uint_as_float(o[j])) + __uint_as_float(q[j]),
                                       (__uint_as_float(r[j + 1]) + __uint_as_float(o[j + 1])) + __uint_as_float(q[j + 1]),
                                       (__uint_as_float(r[j + 2]) + __uint_as_float(o[j + 2])) + __uint_as_float(q[j + 2]),
                                       (__uint_as_float(r[j + 3]) + __uint_as_float(o[j + 3])) + __uint_as_float(q[j + 3]));
          *reinterpret_cast<float4*>(srow + ((((j >> 2) ^ (lane & 7))) << 4)) = v;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(acc_free);    // TMEM accumulators may be overwritten by the next tile
      mbar_arrive(stage_full);  // (release: the staging writes above are visible to the store warps)
    }
  } else if (warp >= 8) {
    // ---- epilogue, second half: staging -> bias / relu6 / mask -> global, overlapping the next tile's main loop ----
    const int w8 = warp - 8, t8 = threadIdx.x - 256;
    int ti = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++ti) {
      int z, m0, n0, k_begin, nkb;
      decode(t, z, m0, n0, k_begin, nkb);
      mbar_wait_sleep(stage_full, (uint32_t)(ti & 1), p.wait_ns);
      if (B_MN && p.colsum != nullptr && m0 == 0) {
        float sum = 0.f;
#pragma unroll
        for (int gq = 0; gq < 16; ++gq) sum += scr[gq * 128 + t8];
        if (n0 + t8 < p.N) p.colsum[(size_t)z * p.N + n0 + t8] = sum;
      }
      const int n = n0 + lane * 4;
      const bool n_ok = n < p.N;  // N % 4 == 0
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n_ok && (p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_RELU6)) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      const unsigned char* sbuf = staging + (size_t)(lane >> 3) * TC_TILE_BYTES;
      const int rows_here = min(TC_BM, p.M - m0);
      float* crow = p.C + ((size_t)z * p.M + m0) * p.ldc + n;
      if (n_ok) {
#pragma unroll 1
        for (int r0 = w8; r0 < rows_here; r0 += 32) {
          float4 h[8];
          if (p.epi == TC_EPI_MASK6) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int row = r0 + 4 * u;
              h[u] = row < rows_here ? __ldg(reinterpret_cast<const float4*>(p.Hm + (size_t)(m0 + row) * p.ldh + n))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int row = r0 + 4 * u;
            if (row >= rows_here) break;
            float4 v = *reinterpret_cast<const float4*>(sbuf + (size_t)row * 128 + ((((lane & 7) ^ (row & 7))) << 4));
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            if (p.epi == TC_EPI_BIAS_RELU6) {
              v.x = fminf(fmaxf(v.x, 0.f), 6.f); v.y = fminf(fmaxf(v.y, 0.f), 6.f);
              v.z = fminf(fmaxf(v.z, 0.f), 6.f); v.w = fminf(fmaxf(v.w, 0.f), 6.f);
            }
            if (p.epi == TC_EPI_MASK6) {
              v.x = (h[u].x > 0.f && h[u].x < 6.f) ? v.x : 0.f; v.y = (h[u].y > 0.f && h[u].y < 6.f) ? v.y : 0.f;
              v.z = (h[u].z > 0.f && h[u].z < 6.f) ? v.z : 0.f; v.w = (h[u].w > 0.f && h[u].w < 6.f) ? v.w : 0.f;
            }
            *reinterpret_cast<float4*>(crow + (size_t)row * p.ldc) = v;
          }
        }
      }
      mbar_arrive(stage_free);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-PAIR form (round 2, K-major A only: forward and input gradient): two CTAs of one cluster -- the two SMs of a TPC --
// compute a 256 x 128 output tile with tcgen05.mma.cta_group::2.  CTA r of the pair owns rows [m0 + 128 r, + 128): its
// own A tile feeds the hi products straight from shared memory and, through its own splitter warps, the lo products from
// its OWN tensor memory; its accumulators (128 lanes x 384 columns) live in its own tensor memory, and its shared memory
// holds only HALF of the B tile (64 of the 128 N-rows, hi and lo).
// The one-CTA kernel is bound by shared-memory bandwidth, not by the tensor pipe: per 32-wide k-block an SM absorbs
// 48 KB of TMA writes, 16 KB of splitter reads and 12 x 4 KB of B-operand reads by the MMAs = 112 KB against 128 B/clk,
// i.e. >= 875 clk where the 12 MMAs need ~800.  The pair halves the B traffic per SM (72 KB, ~560 clk).
// Only the leader (cluster rank 0) issues MMAs.  Barriers: full / empty / acc_full / stage_* are per CTA (empty and
// acc_full are signalled in BOTH CTAs by one multicast tcgen05.commit); conv and acc_free live in the LEADER and count
// the 2 x 128 splitter threads of both CTAs (remote arrive, release / acquire at cluster scope).
// ------------------------------------------------------------------------------------------------
constexpr int TP_STAGES = 4;
static_assert(TP_STAGES * 32 <= 512 - (int)TC_TMEM_A, "one 32-column a_lo slot per ring stage next to the three accumulators");
constexpr int TP_HALF_BYTES = TC_TILE_BYTES / 2;                       // 64 N-rows x 32 floats
constexpr int TP_STAGE_BYTES = TC_TILE_BYTES + 2 * TP_HALF_BYTES;      // A, B_hi half, B_lo half
constexpr int TP_RING_BYTES = TP_STAGES * TP_STAGE_BYTES;
constexpr int TP_SMEM_BYTES = TP_RING_BYTES + TCP_STAGING_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Remote arrive with the DEFAULT semantics (release at CTA scope), as CUTLASS's ClusterBarrier does.  What the leader's
// MMAs read from the peer is (a) tensor memory written by tcgen05.st -- ordered by tcgen05.wait::st +
// tcgen05.fence::before_thread_sync on this side and fence::after_thread_sync on the leader's -- and (b) shared memory
// written by TMA, complete (complete_tx on this CTA's `full` barrier) before this thread arrives; neither sits behind a
// cache.  A `.release.cluster` arrive compiles to MEMBAR.ALL.GPU + ERRBAR per thread per k-block: 13 % of all stall
// samples of the first pair profile, on the critical splitter -> MMA path (1972 instead of ~1000 clk per k-block).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// M = 256 over the pair: A rows [0,128) from the leader's tensor memory, [128,256) from the peer's (same address),
// B = N x 8 with N/2 rows from each CTA's shared memory (same offset), D rows likewise split over the two tensor memories
__device__ __forceinline__ void umma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on `bar` (same offset) in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

// both operands from shared memory: A = this CTA's 128 rows (K-major tile as TMA landed it), B as above
__device__ __forceinline__ void umma_tf32_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

template <bool B_MN>
__global__ void __launch_bounds__(TCP_THREADS, 1) tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                      const __grid_constant__ CUtensorMap mapB,
                                                                      const __grid_constant__ CUtensorMap mapBlo, const TcParams p,
                                                                      const int tiles_n, const int num_tiles) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;  // (the dynamic segment starts at the same offset in both CTAs)
  unsigned char* gbase = smem_dyn + (base - raw);
  unsigned char* staging = gbase + TP_RING_BYTES;
  const uint32_t bar0 = base + TP_RING_BYTES + TCP_STAGING_BYTES;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto conv = [&](int s) { return bar0 + 8 * (TP_STAGES + s); };
  auto empty = [&](int s) { return bar0 + 8 * (2 * TP_STAGES + s); };
  const uint32_t acc_full = bar0 + 8 * 3 * TP_STAGES, acc_free = acc_full + 8, stage_full = acc_full + 16,
                 stage_free = acc_full + 24;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(gbase + TP_RING_BYTES + TCP_STAGING_BYTES + 8 * (3 * TP_STAGES + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TP_STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 256);  // (used in the leader only)
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 256);   // (leader only)
    mbar_init(stage_full, 128);
    mbar_init(stage_free, 128);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    if (p.b_lo_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBlo)) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  auto decode = [&](int t, int& m0, int& n0) {
    const int mb = t / tiles_n;
    m0 = mb * 2 * TC_BM + (int)crank * TC_BM;  // this CTA's 128 rows
    n0 = (t - mb * tiles_n) * TC_BN;
  };
  const int nkb = (p.K + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        int m0, n0;
        decode(t, m0, n0);
        const int nh = n0 + (int)crank * (TC_BN / 2);  // this CTA's half of the B tile
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % TP_STAGES;
          mbar_wait_sleep(empty(s), (uint32_t)(((g / TP_STAGES) & 1) ^ 1), p.wait_ns);
          mbar_expect_tx(full(s), TC_TILE_BYTES + (p.b_lo_tma ? 2 : 1) * TP_HALF_BYTES);
          const uint32_t a_dst = base + s * TP_STAGE_BYTES, b_dst = a_dst + TC_TILE_BYTES, l_dst = b_dst + TP_HALF_BYTES;
          const int k0 = kb * TC_BK;
          tma_load_2d(a_dst, &mapA, k0, m0, full(s));
          if (B_MN) {
#pragma unroll
            for (int q = 0; q < TC_BN / 64; ++q) tma_load_2d(b_dst + q * 4096, &mapB, nh + 32 * q, k0, full(s));
          } else {
            tma_load_2d(b_dst, &mapB, k0, nh, full(s));  // (box of 64 rows)
          }
          if (p.b_lo_tma) {
            if (B_MN) {
#pragma unroll
              for (int q = 0; q < TC_BN / 64; ++q) tma_load_2d(l_dst + q * 4096, &mapBlo, nh + 32 * q, k0, full(s));
            } else {
              tma_load_2d(l_dst, &mapBlo, k0, nh, full(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(TC_BN >> 3) << 17) |
                             ((uint32_t)((2 * TC_BM) >> 4) << 24);
      int g = 0, ti = 0;
      for (int t = pair; t < num_tiles; t += npairs, ++ti) {
        if (ti > 0) mbar_wait(acc_free, (uint32_t)((ti - 1) & 1));  // both CTAs read the previous accumulators out
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % TP_STAGES;
          mbar_wait(conv(s), (uint32_t)((g / TP_STAGES) & 1));  // both A tiles are in tensor memory, both B halves landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_sm = base + s * TP_STAGE_BYTES, b_hi = a_sm + TC_TILE_BYTES, b_lo = b_hi + TP_HALF_BYTES;
          const uint32_t a_lo = tmem + TC_TMEM_A + (uint32_t)(s * 32);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            // a_hi is the raw A tile where TMA put it (the tensor core drops the low 13 mantissa bits itself); only a_lo
            // goes through the splitter into tensor memory -- one 32-column slot PER RING STAGE, so the splitter runs up to
            // TP_STAGES - 1 k-blocks ahead of the MMAs.  (With hi AND lo in tensor memory only two slots fit next to the
            // three accumulators, and the commit -> tcgen05.st -> arrive -> issue round trip, ~1500 clk, bounded the
            // k-block period at (800 + 1500) / 2 clk in both the one-CTA and the first pair kernel.)
            const uint64_t da_hi = umma_desc(a_sm + ks * 32);
            const uint64_t db_hi = B_MN ? umma_desc_mn(b_hi + ks * 1024) : umma_desc(b_hi + ks * 32);
            const uint64_t db_lo = B_MN ? umma_desc_mn(b_lo + ks * 1024) : umma_desc(b_lo + ks * 32);
            const uint32_t dmain = tmem + (uint32_t)((kb & 1) * TC_BN);
            umma_tf32_ss_pair(dmain, da_hi, db_hi, idesc, (kb >= 2 || ks != 0) ? 1u : 0u);
            umma_tf32_ts_pair(tmem + 2 * TC_BN, a_lo + ks * 8, db_hi, idesc, (kb | ks) != 0);
            umma_tf32_ss_pair(tmem + 2 * TC_BN, da_hi, db_lo, idesc, 1u);
          }
          umma_commit_pair(empty(s));
        }
        umma_commit_pair(acc_full);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int t128 = threadIdx.x - 128;
    const int wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t conv_leader0 = mapa_cluster(conv(0), 0), acc_free_leader = mapa_cluster(acc_free, 0);
    int g = 0, ti = 0;
    for (int t = pair; t < num_tiles; t += npairs, ++ti) {
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const int s = g % TP_STAGES;
        mbar_wait(full(s), (uint32_t)((g / TP_STAGES) & 1));
        const unsigned char* sa = gbase + (size_t)s * TP_STAGE_BYTES;
        uint32_t xl[32];
        const unsigned char* pa = sa + t128 * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(pa + ((c ^ (t128 & 7)) << 4));
          xl[4 * c] = v.x; xl[4 * c + 1] = v.y; xl[4 * c + 2] = v.z; xl[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < 32; ++r)
          xl[r] = __float_as_uint(__uint_as_float(xl[r]) - __uint_as_float(xl[r] & 0xffffe000u));
        if (!p.b_lo_tma) {  // this CTA's half of B_lo, elementwise in the landed (swizzled) layout
          const float4* hi = reinterpret_cast<const float4*>(sa + TC_TILE_BYTES);
          float4* lo = reinterpret_cast<float4*>(const_cast<unsigned char*>(sa) + TC_TILE_BYTES + TP_HALF_BYTES);
#pragma unroll
          for (int j = 0; j < TP_HALF_BYTES / 16 / 128; ++j) {
            const int i = t128 + 128 * j;
            const float4 x = hi[i];
            float4 l;
            l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            lo[i] = l;
          }
        }
        // (slot s was last read by the MMAs of k-block g - TP_STAGES; the TMA warp waited for their commit before it
        //  refilled stage s, and this thread saw that refill complete on full(s))
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_st32(tmem + lane_base + TC_TMEM_A + (uint32_t)(s * 32), xl);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (!p.b_lo_tma) fence_async_smem();  // (the B_lo half written above is read by the MMAs through the async proxy)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive_cluster(conv_leader0 + 8 * s);
      }
      // ---- epilogue, first half: own rows TMEM -> staging; then the accumulators are free again -------------------
      mbar_wait(acc_full, (uint32_t)(ti & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (ti > 0) mbar_wait(stage_free, (uint32_t)((ti - 1) & 1));  // the store warps are done with the previous tile
#pragma unroll 1
      for (int c0 = 0; c0 < TC_BN; c0 += 32) {
        uint32_t r[32], q[32], o[32];
        const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
              "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
              "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
              "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
            : "r"(taddr + (uint32_t)(2 * TC_BN)));
        if (nkb > 1) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]),
                "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15]), "=r"(o[16]),
                "=r"(o[17]), "=r"(o[18]), "=r"(o[19]), "=r"(o[20]), "=r"(o[21]), "=r"(o[22]), "=r"(o[23]), "=r"(o[24]),
                "=r"(o[25]), "=r"(o[26]), "=r"(o[27]), "=r"(o[28]), "=r"(o[29]), "=r"(o[30]), "=r"(o[31])
              : "r"(taddr + (uint32_t)TC_BN));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0u;
        }
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        unsigned char* srow = staging + (size_t)(c0 >> 5) * TC_TILE_BYTES + (size_t)(wq * 32 + lane) * 128;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 v = make_float4((__uint_as_float(r[j]) + __uint_as_float(o[j])) + __uint_as_float(q[j]),
                                       (__uint_as_float(r[j + 1]) + __uint_as_float(o[j + 1])) + __uint_as_float(q[j + 1]),
                                       (__uint_as_float(r[j + 2]) + __uint_as_float(o[j + 2])) + __uint_as_float(q[j + 2]),
                                       (__uint_as_float(r[j + 3]) + __uint_as_float(o[j + 3])) + __uint_as_float(q[j + 3]));
          *reinterpret_cast<float4*>(srow + ((((j >> 2) ^ (lane & 7))) << 4)) = v;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive_cluster(acc_free_leader);  // this CTA's accumulators may be overwritten by the next tile
      mbar_arrive(stage_full);               // (release: the staging writes above are visible to the store warps)
    }
  } else if (warp >= 8) {
    // ---- epilogue, second half: staging -> bias / relu6 / mask -> global, overlapping the next tile's main loop ----
    const int w8 = warp - 8;
    int ti = 0;
    for (int t = pair; t < num_tiles; t += npairs, ++ti) {
      int m0, n0;
      decode(t, m0, n0);
      mbar_wait_sleep(stage_full, (uint32_t)(ti & 1), p.wait_ns);
      const int n = n0 + lane * 4;
      const bool n_ok = n < p.N;  // N % 4 == 0
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n_ok && (p.epi == TC_EPI_BIAS || p.epi == TC_EPI_BIAS_RELU6)) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      const unsigned char* sbuf = staging + (size_t)(lane >> 3) * TC_TILE_BYTES;
      const int rows_here = min(TC_BM, p.M - m0);  // (<= 0 for the peer CTA of a ragged last row block)
      if (n_ok && rows_here > 0) {
        float* crow = p.C + (size_t)m0 * p.ldc + n;
#pragma unroll 1
        for (int r0 = w8; r0 < rows_here; r0 += 32) {
          float4 h[8];
          if (p.epi == TC_EPI_MASK6) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int row = r0 + 4 * u;
              h[u] = row < rows_here ? __ldg(reinterpret_cast<const float4*>(p.Hm + (size_t)(m0 + row) * p.ldh + n))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int row = r0 + 4 * u;
            if (row >= rows_here) break;
            float4 v = *reinterpret_cast<const float4*>(sbuf + (size_t)row * 128 + ((((lane & 7) ^ (row & 7))) << 4));
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            if (p.epi == TC_EPI_BIAS_RELU6) {
              v.x = fminf(fmaxf(v.x, 0.f), 6.f); v.y = fminf(fmaxf(v.y, 0.f), 6.f);
              v.z = fminf(fmaxf(v.z, 0.f), 6.f); v.w = fminf(fmaxf(v.w, 0.f), 6.f);
            }
            if (p.epi == TC_EPI_MASK6) {
              v.x = (h[u].x > 0.f && h[u].x < 6.f) ? v.x : 0.f; v.y = (h[u].y > 0.f && h[u].y < 6.f) ? v.y : 0.f;
              v.z = (h[u].z > 0.f && h[u].z < 6.f) ? v.z : 0.f; v.w = (h[u].w > 0.f && h[u].w < 6.f) ? v.w : 0.f;
            }
            *reinterpret_cast<float4*>(crow + (size_t)row * p.ldc) = v;
          }
        }
      }
      mbar_arrive(stage_free);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared and tensor memory: nobody leaves before both are done
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return q == cudaDriverEntryPointSuccess ? reinterpret_cast<EncodeTiledFn>(f) : nullptr;
  }();
  return fn;
}

// K-major: [rows, K] row-major fp32 -> boxes of (box_rows x 32 floats along K);
// MN-major: [K, rows] row-major fp32 -> boxes of (32 k-rows x 32 floats along M/N).  128-byte swizzle (16-byte
// chunks for K-major, 32-byte chunks for MN-major), OOB reads return 0.
static int make_map(CUtensorMap* map, const float* ptr, int rows, int K, int ld, int box_rows, bool mn_major = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return PFPN_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)(mn_major ? rows : K), (cuuint64_t)(mn_major ? K : rows)};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)(mn_major ? TC_BK : box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PFPN_OK : PFPN_ERR_UNSUPPORTED;
}

template <bool A_MN, bool B_MN>
static int tc_launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapBlo, const TcParams& p, dim3 grid,
                     cudaStream_t st) {
  static int attr_dev = -1;  // per instantiation; the attribute is per device
  int dev = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  if (dev != attr_dev) {
    PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)tc_gemm_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TC_SMEM_BYTES));
    attr_dev = dev;
  }
  static const int persist = []() {
    const char* e = getenv("PFPN_TC_PERSISTENT");
    return e ? atoi(e) : 1;
  }();
  const int num_tiles = (int)(grid.x * grid.y * grid.z);
  // Measured at M = 65536 (ms, persistent vs one-tile-per-CTA): forward K=200 0.180 vs 0.235, N=1260 K=512 0.415 vs 0.464,
  // input gradient N=1024 K=512 0.356 vs 0.425, K=1024 0.326 vs 0.325; the split-K weight gradients (MN-major A) are 3-6 %
  // slower persistent (0.443 vs 0.427), so they keep the one-tile form.
  if (persist && num_tiles > 1 && !A_MN) {
    static int attr_dev_p = -1;
    static int sms = 0;
    if (dev != attr_dev_p) {
      PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)tc_gemm_persist_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TCP_SMEM_BYTES));
      PFPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      attr_dev_p = dev;
    }
    const int per_cta = (num_tiles + sms - 1) / sms;
    const int ctas = (num_tiles + per_cta - 1) / per_cta;  // (balanced: see tc_launch_pair)
    tc_gemm_persist_kernel<A_MN, B_MN><<<ctas, TCP_THREADS, TCP_SMEM_BYTES, st>>>(mapA, mapB, mapBlo, p, (int)grid.y, (int)grid.x,
                                                                                  num_tiles);
    PFPN_CUDA_OK(cudaGetLastError());
    return PFPN_OK;
  }
  tc_gemm_kernel<A_MN, B_MN><<<grid, 256, TC_SMEM_BYTES, st>>>(mapA, mapB, mapBlo, p);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

// CTA-pair launch (K-major A): clusters of two CTAs, as many pairs as the device keeps resident at once (<= SMs / 2)
template <bool B_MN>
static int tc_launch_pair(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapBlo, const TcParams& p, int tiles_n,
                          int num_tiles, cudaStream_t st) {
  static int attr_dev = -1, max_pairs = 0;
  int dev = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cfg.blockDim = dim3(TCP_THREADS);
  cfg.dynamicSmemBytes = TP_SMEM_BYTES;
  cfg.stream = st;
  if (dev != attr_dev) {
    PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)tc_gemm_pair_kernel<B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES));
    int sms = 0;
    PFPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cfg.gridDim = dim3((unsigned)(sms & ~1));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)tc_gemm_pair_kernel<B_MN>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = sms / 2;
    }
    max_pairs = n < sms / 2 ? n : sms / 2;
    attr_dev = dev;
  }
  // balanced grid: every pair walks the same number of tiles (ceil), and no more pairs are launched than that needs --
  // the SMs left over serve the kernels of the other streams (critic branch, weight gradients) instead of idling through
  // a ragged last wave (8192 rows x N=1024: 256 tiles = 64 pairs x 4, not 74 pairs x 3.46)
  const int per_pair = (num_tiles + max_pairs - 1) / max_pairs;
  const int pairs = (num_tiles + per_pair - 1) / per_pair;
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  PFPN_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_gemm_pair_kernel<B_MN>, mapA, mapB, mapBlo, p, tiles_n, num_tiles));
  return PFPN_OK;
}

static int tc_gemm_common(const float* A, int lda, const float* Bm, const float* Blo, int ldb, bool b_mn, float* Cout, int ldc,
                          const float* bias, const float* Hm, int ldh, int M, int N, int K, int epi, cudaStream_t st) {
  if (!A || !Bm || !Cout || M < 0 || N <= 0 || K <= 0 || epi < 0 || epi > 3) return PFPN_ERR_ARG;
  if ((epi == TC_EPI_BIAS || epi == TC_EPI_BIAS_RELU6) && !bias) return PFPN_ERR_ARG;
  if (epi == TC_EPI_MASK6 && !Hm) return PFPN_ERR_ARG;
  if (M == 0) return PFPN_OK;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if ((lda & 3) || (ldb & 3) || (ldc & 3) || (N & 3) || (K & 3) || !al16(A) || !al16(Bm) || !al16(Cout) || !al16(Blo))
    return PFPN_ERR_ALIGN;
  // PFPN_TC_PAIR (default 1): two-SM MMAs (tcgen05 cta_group::2) once there are at least two row blocks
  static const int pair_on = []() {
    const char* e = getenv("PFPN_TC_PAIR");
    return e ? atoi(e) : 1;
  }();
  const bool use_pair = pair_on && M > TC_BM;
  CUtensorMap mapA, mapB, mapBlo;
  int rc = make_map(&mapA, A, M, K, lda, TC_BM);
  if (rc != PFPN_OK) return rc;
  rc = make_map(&mapB, Bm, N, K, ldb, use_pair ? TC_BN / 2 : TC_BN, b_mn);
  if (rc != PFPN_OK) return rc;
  rc = make_map(&mapBlo, Blo ? Blo : Bm, N, K, ldb, use_pair ? TC_BN / 2 : TC_BN, b_mn);
  if (rc != PFPN_OK) return rc;
  TcParams p{Cout, bias, Hm, M, N, K, ldc, ldh, epi, (K + TC_BK - 1) / TC_BK * TC_BK, nullptr, pfpn_wait_ns(64u), Blo ? 1 : 0};
  if (use_pair) {
    const int tiles_n = (N + TC_BN - 1) / TC_BN, tiles_m2 = (M + 2 * TC_BM - 1) / (2 * TC_BM);
    return b_mn ? tc_launch_pair<true>(mapA, mapB, mapBlo, p, tiles_n, tiles_n * tiles_m2, st)
                : tc_launch_pair<false>(mapA, mapB, mapBlo, p, tiles_n, tiles_n * tiles_m2, st);
  }
  dim3 grid((N + TC_BN - 1) / TC_BN, (M + TC_BM - 1) / TC_BM);
  return b_mn ? tc_launch<false, true>(mapA, mapB, mapBlo, p, grid, st) : tc_launch<false, false>(mapA, mapB, mapBlo, p, grid, st);
}

}  // namespace pfpn

using namespace pfpn;

// C[M,N] = epi(A[M,K] * Bt[N,K]^T): the tensor-core twin of pfpn_mlp_linear_bwd_input (Bt = W as stored, [in, out]
// seen as [N = in, K = out]).  epi: 0 none, 1 +bias, 2 relu6(+bias), 3 Relu6Grad mask by Hm.
extern "C" int pfpn_tc_gemm_nt(const float* A, int32_t lda, const float* Bt, int32_t ldb, float* Cout, int32_t ldc,
                               const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K,
                               int32_t epi, pfpn_stream_t stream_) {
  return tc_gemm_common(A, lda, Bt, nullptr, ldb, false, Cout, ldc, bias, Hm, ldh, M, N, K, epi, reinterpret_cast<cudaStream_t>(stream_));
}
// Same with the low part of the B operand supplied (Bt_lo = Bt - tf32(Bt), pfpn_split_lo; same layout and ldb): the weights
// are constant within an optimizer step, so their split is done once per step instead of once per tile per GEMM.
extern "C" int pfpn_tc_gemm_nt_lo(const float* A, int32_t lda, const float* Bt, const float* Bt_lo, int32_t ldb, float* Cout,
                                  int32_t ldc, const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K,
                                  int32_t epi, pfpn_stream_t stream_) {
  if (!Bt_lo) return PFPN_ERR_ARG;
  return tc_gemm_common(A, lda, Bt, Bt_lo, ldb, false, Cout, ldc, bias, Hm, ldh, M, N, K, epi, reinterpret_cast<cudaStream_t>(stream_));
}

// C[M,N] = epi(A[M,K] * B[K,N]): the tensor-core twin of pfpn_mlp_linear_fwd with W as stored ([in, out] row-major,
// read as an MN-major operand -- no transposed copy of the weights).
extern "C" int pfpn_tc_gemm_nn(const float* A, int32_t lda, const float* B, int32_t ldb, float* Cout, int32_t ldc,
                               const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K,
                               int32_t epi, pfpn_stream_t stream_) {
  return tc_gemm_common(A, lda, B, nullptr, ldb, true, Cout, ldc, bias, Hm, ldh, M, N, K, epi, reinterpret_cast<cudaStream_t>(stream_));
}
extern "C" int pfpn_tc_gemm_nn_lo(const float* A, int32_t lda, const float* B, const float* B_lo, int32_t ldb, float* Cout,
                                  int32_t ldc, const float* bias, const float* Hm, int32_t ldh, int32_t M, int32_t N, int32_t K,
                                  int32_t epi, pfpn_stream_t stream_) {
  if (!B_lo) return PFPN_ERR_ARG;
  return tc_gemm_common(A, lda, B, B_lo, ldb, true, Cout, ldc, bias, Hm, ldh, M, N, K, epi, reinterpret_cast<cudaStream_t>(stream_));
}

// lo[i] = x[i] - tf32(x[i]) (tf32 = the top 19 bits; exact in fp32): the part of an fp32 operand the tensor core drops
namespace pfpn {
__global__ void split_lo_kernel(const float4* __restrict__ x, float4* __restrict__ lo, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 l;
    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    lo[i] = l;
  }
}
}  // namespace pfpn
extern "C" int pfpn_split_lo(const float* x, float* lo, size_t n, pfpn_stream_t stream_) {
  if (!x || !lo || (n & 3)) return PFPN_ERR_ARG;
  if (n == 0) return PFPN_OK;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lo)) & 15u) return PFPN_ERR_ALIGN;
  size_t grid = (n / 4 + 255) / 256;
  if (grid > 592) grid = 592;
  pfpn::split_lo_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(lo), n / 4);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores: dW[K,N] = X[M,K]^T dY[M,N], X and dY as stored (row-major,
// batch-major): both are MN-major operands of the GEMM whose reduction axis is the batch.  The
// launch is split-K in chunks of 2048 rows (<= 128 accumulations per TMEM accumulator, see above)
// and the partial tiles are summed in fp32 in a fixed order.
// ------------------------------------------------------------------------------------------------
namespace pfpn {
constexpr int TC_WGRAD_CHUNK = 2048;
// batch rows per split: 2048 at scale (<= 128 accumulations per TMEM accumulator); halved while the launch
// would leave SMs idle (small per-GPU minibatches of the strong-scaled DPPO update), never below 256
static int wgrad_chunk(int M, int K, int N) {
  const long long tiles = (long long)((N + TC_BN - 1) / TC_BN) * ((K + TC_BM - 1) / TC_BM);
  int chunk = TC_WGRAD_CHUNK;
  while (chunk > 256 && tiles * ((M + chunk - 1) / chunk) < 148) chunk >>= 1;
  return chunk;
}
__global__ void tc_splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, size_t n4, int splits,
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
}  // namespace pfpn

extern "C" int pfpn_tc_wgrad_workspace_bytes(int32_t M, int32_t K, int32_t N, size_t* bytes) {
  if (!bytes || M <= 0 || K <= 0 || N <= 0) return PFPN_ERR_ARG;
  const int chunk = wgrad_chunk(M, K, N);
  const size_t splits = ((size_t)M + chunk - 1) / chunk;
  *bytes = splits * ((size_t)K * N + N) * sizeof(float) + 256;
  return PFPN_OK;
}

extern "C" int pfpn_tc_linear_bwd_weight(const float* X, int32_t ldx, const float* dY, int32_t ldy, float* dW, float* db,
                                         int32_t M, int32_t K, int32_t N, void* workspace, size_t workspace_bytes,
                                         pfpn_stream_t stream_) {
  if (!X || !dY || !dW || M <= 0 || K <= 0 || N <= 0) return PFPN_ERR_ARG;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if ((ldx & 3) || (ldy & 3) || (N & 3) || !al16(X) || !al16(dY) || !al16(dW) || !al16(db) || !al16(workspace)) return PFPN_ERR_ALIGN;
  const int chunk = wgrad_chunk(M, K, N);
  const int splits = (M + chunk - 1) / chunk;
  size_t need;
  pfpn_tc_wgrad_workspace_bytes(M, K, N, &need);
  if ((splits > 1 || db) && (!workspace || workspace_bytes < need)) return PFPN_ERR_WORKSPACE;
  CUtensorMap mapA, mapB;
  int rc = make_map(&mapA, X, K, M, ldx, TC_BM, true);  // GEMM rows = K_in, reduction axis = batch
  if (rc != PFPN_OK) return rc;
  rc = make_map(&mapB, dY, N, M, ldy, TC_BN, true);
  if (rc != PFPN_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  float* out = splits > 1 ? reinterpret_cast<float*>(workspace) : dW;
  float* cs_part = db ? reinterpret_cast<float*>(workspace) + (size_t)splits * K * N : nullptr;
  TcParams p{out, nullptr, nullptr, K, N, M, N, 0, TC_EPI_NONE, chunk, cs_part, pfpn_wait_ns(64u), 0};
  dim3 grid((N + TC_BN - 1) / TC_BN, (K + TC_BM - 1) / TC_BM, splits);
  rc = tc_launch<true, true>(mapA, mapB, mapB, p, grid, st);
  if (rc != PFPN_OK) return rc;
  if (splits > 1) {
    const size_t n4 = (size_t)K * N / 4;
    tc_splitk_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(out, dW, n4, splits, (size_t)K * N);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  if (db) {  // bias gradient: the per-split column sums of dY, combined in split order
    tc_splitk_reduce_kernel<<<(unsigned)((N / 4 + 255) / 256), 256, 0, st>>>(cs_part, db, (size_t)N / 4, splits, (size_t)N);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  return PFPN_OK;
}
