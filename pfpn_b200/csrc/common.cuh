// Shared device helpers for the pfpn_b200 kernels (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/pfpn_b200.h"

namespace pfpn {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kHalfLog2Pi = 0.9189385175704956f;  // [graph] Normal/prob_1 const
constexpr int kNumSmFallback = 148;

#define PFPN_CUDA_OK(expr)                      \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

// ---- packed fp32x2 math (Blackwell FFMA2 / FADD2 / FMUL2) ------------------
__device__ __forceinline__ unsigned long long& as_u64(float2& v) {
  return reinterpret_cast<unsigned long long&>(v);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(as_u64(d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// ---- mbarrier + 1-D bulk async copies (TMA, SASS: UBLKCP / SYNCS) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
// same, backing off with nanosleep between polls (ns == 0: plain spin).  For warps that wait for most of a tile (a TMA
// producer waiting for its consumers): a hot try_wait loop issues ~3 instructions every ~12 cycles and takes those
// issue slots from the compute warps of the same SM sub-partition -- 10 % of all instructions of K3f in its first profile.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (ns) __nanosleep(ns);
  }
}
// PFPN_WAIT_NS overrides the back-off of the long waits (host side, read once); unset -> the caller's default
inline uint32_t pfpn_wait_ns(uint32_t dflt) {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("PFPN_WAIT_NS");
    v = e ? atoi(e) : -1;
  }
  return v >= 0 ? (uint32_t)v : dflt;
}
// same, with a suspend-time hint (ns): the hardware may park the thread that long before try_wait returns false,
// instead of the warp spinning through SYNCS / BRA / YIELD issue slots
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// 4-byte LDGSTS and "arrive on this mbarrier when my cp.asyncs have landed"
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- Philox4x32-R (counter-based RNG; Salmon et al. 2011) -----------------
// R = 10 is the standard variant; R = 7 is the fewest rounds that still passes BigCrush in the paper ("Philox4x32-7") and
// is what the RNG-bound SAC head kernels (K3 / K3f: ~40 % of their instructions were Philox) use.
template <int ROUNDS>
struct PhiloxR {
  uint32_t key[2];
  __device__ __forceinline__ PhiloxR(uint64_t seed) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
  }
  __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi,
             c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;  // one IMAD.WIDE each
      const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
      c0 = hi1 ^ c1 ^ k0;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ k1;
      c3 = lo0;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
using Philox = PhiloxR<10>;
using Philox7 = PhiloxR<7>;
// uniform in (0,1): (x + 0.5) * 2^-32 never returns 0 or 1 in fp32? (rounding may give 1.0f) -> clamp
__device__ __forceinline__ float u32_to_unit_open(uint32_t x) {
  float u = ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // 24-bit, strictly inside (0,1)
  return u;
}
// One standard normal from two 24-bit uniforms (Box-Muller, cosine branch) on the MUFU pipe: sqrt.approx(-2 ln u1)
// cos.approx(2 pi u2).  Shared by K2 (csrc/sample.cu) and K2f (csrc/rollout.cu) so that both produce the same bits for the
// same Philox block; absolute error ~1e-6, far below the resolution of the 24-bit uniforms it is fed.
__device__ __forceinline__ float normal_from_bits(uint32_t x, uint32_t y) {
  const float u1 = u32_to_unit_open(x), u2 = u32_to_unit_open(y);
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.f * 0.6931471805599453f * l));
  return r * __cosf(6.283185307179586f * u2);
}
__device__ __forceinline__ double u64_to_unit_double(uint32_t hi, uint32_t lo) {
  // 53-bit mantissa in [0,1)
  uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace pfpn
