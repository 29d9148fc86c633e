// K2f -- the rollout side of the head in ONE pass over logits[B, A, P]: particle sampling (TF Multinomial CPU semantics),
// the action, its mixture log_prob, the categorical entropy and the running activity statistics.
//
// Reference: what ClipPPONetwork.run executes per environment step (ppo.py:56-62 -> actor_critic.py:368-380):
//   policy.sample(1)                 networks/utils.py:187-194  (Categorical.sample -> Multinomial; Normal.sample; gather)
//   policy.log_prob(action)          networks/utils.py:108-144
//   running_update_ops               networks/actor_critic/a2c.py:346-365  (max_active, sum_active over the batch)
// Round 1 ran three kernels over the same logits (K2 sample_kernel 636 us, K1 forward 83 us, K4 stats_kernel 364 us at
// B = 65536: 0.08 / 0.63 / 0.14 of the HBM roofline); this kernel reads them once (4AP + 8A + 8 bytes per state).
//
// Skeleton of K1 / K3f: persistent CTAs, tiles of SLOTS states fetched by one bulk async copy (TMA) into a shared-memory
// ring, a producer warp that also sums the per-row log p / entropy of a state in fixed order; a mixture row is owned
// by 2 adjacent lanes with the bank-conflict-free split ownership of K1 (lane c: particles 16c .. 16c+15, then 32 + c,
// then 34 for c = 0); per-(a,k) statistics accumulate in registers (a thread's particle set never changes) and are
// combined across slots / CTAs in a fixed order (max is exact, the sum order is fixed) -- no atomics.
//
// Sampling.  The TF-1.14 CPU Multinomial functor builds an fp64 running CDF of exp(double(logit) - max) and returns
// upper_bound(cdf, u * total).  As in K2 the interval is located with an fp32 CDF whose error is bounded by
// delta * total and accepted only if u * total keeps a (2 delta + 2^-23) * total margin from EVERY fp32 CDF value;
// otherwise (~0.1 % of rows) lane 0 of the row redoes it literally in fp64.  The index is the number of CDF values
// <= u * total, counted by both lanes over their own particles (no search).  Indices are therefore exactly those of the
// fp64 algorithm, with caller-supplied uniforms bit-exact against the oracle.
#include <string.h>

#include "common.cuh"

namespace pfpn {

constexpr int kRoMaxCtas = 148 * 2;
// value a non-finite logit (and the unused 36th half-slot) is replaced by: far below any logit that can matter, yet
// d = kRoExcluded - max stays finite, so no clamp is needed before 0 * d (rows whose maximum is not above it take the
// literal fp64 path, which reads the raw logits)
constexpr float kRoExcluded = -1e30f;

struct RolloutK {
  pfpn_rollout_args a;
  float* part;  // [grid][2][A*P]: per-CTA max / sum of the probabilities
  int num_tiles;
  uint32_t wait_ns;  // producer back-off while the compute threads work on a tile
};

// The literal fp64 algorithm for one row (fallback; also the semantic definition), executed by a whole WARP for one row
// at a time in warp-uniform control flow: every lane evaluates exp(double(logit) - max) for its particles (the expensive
// part), the values go through a shared-memory scratch row, and ONE lane replays TF's sequential fp64 running total and
// upper_bound on them -- so the additions happen in exactly the reference's order.  (A per-lane subroutine call inside
// divergent code ahead of the warp shuffles below deadlocked on sm_100a; this form has no divergent collective at all.)
__device__ __forceinline__ int ro_multinomial_fp64_warp(const float* x, int P, double u, double* scratch, int lane) {
  float m = -3.402823466e38f;
  for (int k = lane; k < P; k += 32) {
    const float v = x[k];
    if (isfinite(v)) m = fmaxf(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  for (int k = lane; k < P; k += 32) {
    const float v = x[k];
    scratch[k] = isfinite(v) ? exp((double)v - (double)m) : 0.0;
  }
  __syncwarp();
  int res = P - 1;  // (u * total == total through rounding: TF would return the undefined class P; clamp as the oracle does)
  if (lane == 0) {
    double total = 0.0;
    for (int k = 0; k < P; ++k) total += scratch[k];
    const double to_find = u * total;
    double run = 0.0;
    for (int k = 0; k < P; ++k) {  // first k with cdf[k] > to_find == std::upper_bound on the running totals
      run += scratch[k];
      if (to_find < run) {
        res = k;
        break;
      }
    }
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, res, 0);
}

template <int SLOTS, int NSTAGE>
__global__ void __launch_bounds__(SLOTS * 72 + 32, 2) rollout_kernel(const RolloutK kp) {
  constexpr int P = 35, A = 36, AP = A * P, LPR = 2, EPL = 18;
  constexpr int NTHR = SLOTS * A * LPR;
  constexpr int TILE_F = SLOTS * AP;
  constexpr int STAGE_BYTES = (TILE_F * 4 + 127) & ~127;
  const int B = kp.a.B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == NTHR / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* tail = smem_raw + (size_t)NSTAGE * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* done_bar = full_bar + NSTAGE;
  float2* rowbuf = reinterpret_cast<float2*>(tail + 16 * NSTAGE);  // [NSTAGE][SLOTS * A] per-row (log p, H)
  float4* cs = reinterpret_cast<float4*>(rowbuf + NSTAGE * SLOTS * A);  // [EPL][A * LPR] per-thread-column constants (one LDS.128)
  float2* musd = reinterpret_cast<float2*>(cs + EPL * A * LPR);         // [AP] {loc, exp(logstd)} of the sampled particle
  double* fb_s = reinterpret_cast<double*>(musd + AP);                  // [compute warps][40] fp64 fallback scratch
  float* red = reinterpret_cast<float*>(smem_raw);  // [SLOTS][2][AP] end-of-kernel combine, aliases the (then idle) stages
  static_assert(SLOTS * 2 * AP * 4 <= NSTAGE * STAGE_BYTES, "the combine tables reuse the stage ring");

  if (is_producer && lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&done_bar[s]), (uint32_t)NTHR);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  int my_tiles = 0;
  if (first_tile < kp.num_tiles) my_tiles = (kp.num_tiles - 1 - first_tile) / tile_step + 1;
  const bool tail_exists = (B % SLOTS) != 0;
  const int tail_tile = kp.num_tiles - 1;
  const float* __restrict__ g_logits = kp.a.logits;
  unsigned char* stage0 = smem_raw;

  if (is_producer) {
    auto issue_load = [&](int it) {
      const int tile = first_tile + it * tile_step;
      if (it >= my_tiles || (tail_exists && tile == tail_tile)) return;
      const int st = it % NSTAGE;
      const uint32_t bar = smem_u32(&full_bar[st]);
      mbar_expect_tx(bar, (uint32_t)(TILE_F * 4));
      bulk_g2s(smem_u32(stage0) + st * STAGE_BYTES, g_logits + (size_t)tile * TILE_F, (uint32_t)(TILE_F * 4), bar);
    };
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < NSTAGE; ++d) issue_load(d);
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % NSTAGE;
      const int tile = first_tile + it * tile_step;
      mbar_wait_sleep(smem_u32(&done_bar[st]), (uint32_t)((it / NSTAGE) & 1), kp.wait_ns);  // every compute thread is done with this stage
      if (lane < SLOTS) {
        const int b = tile * SLOTS + lane;
        if (b < B) {
          const float2* rb = rowbuf + st * SLOTS * A + lane * A;
          float lp = 0.f, en = 0.f;
#pragma unroll 4
          for (int aa = 0; aa < A; ++aa) {
            lp += rb[aa].x;
            en += rb[aa].y;
          }
          kp.a.lp[b] = lp;
          if (kp.a.ent != nullptr) kp.a.ent[b] = en;
        }
      }
      __syncwarp();
      if (lane == 0) issue_load(it + NSTAGE);  // (nothing is written back: the stage is free as soon as it was read)
      __syncwarp();
    }
  } else {
    const int slot = tid / (A * LPR);
    const int rem = tid - slot * (A * LPR);
    const int a = rem >> 1, c = rem & 1;
    // particle of slot i: 16 c + i (i < 16), 32 + c (i = 16), 34 (i = 17, c = 0 only)
    auto kof = [&](int i) -> int { return i < 16 ? 16 * c + i : (i == 16 ? 32 + c : 34); };
    // per-(a,k) constants {1/sigma, -mu/sigma, -(logstd + ln sqrt(2 pi)) log2 e} in shared memory, one column per thread
    // of a state slot (conflict-free); the statistics accumulators stay in registers
    float4* csp = cs + rem;
    constexpr int TPS = A * LPR;
    float vmax[EPL], vsum[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const bool ok = i < 17 || c == 0;
      const int k = ok ? kof(i) : 0;
      const float ls = __ldg(&kp.a.logstd[a * P + k]), mu = __ldg(&kp.a.loc[a * P + k]);
      const float is_ = ok ? expf(-ls) : 0.f;
      if (slot == 0) {
        csp[i * TPS] = make_float4(is_, ok ? -mu * is_ : 0.f, ok ? -(ls + kHalfLog2Pi) * kLog2e : 0.f, 0.f);
        if (ok) musd[a * P + k] = make_float2(mu, expf(ls));  // (expf as K2 evaluates it: the same action bits)
      }
      vmax[i] = 0.f;
      vsum[i] = 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NTHR) : "memory");  // constants visible to every slot
    const Philox7 rng(kp.a.seed);  // (as K2)
    const float delta = 4e-6f + (float)P * 6e-8f;  // error budget of the fp32 CDF relative to the total (as K2)

    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % NSTAGE;
      const int tile = first_tile + it * tile_step;
      float* sbuf = reinterpret_cast<float*>(stage0 + (size_t)st * STAGE_BYTES);
      const bool is_tail = tail_exists && tile == tail_tile;
      const int b0 = tile * SLOTS;
      if (!is_tail) {
        mbar_wait(smem_u32(&full_bar[st]), (uint32_t)((it / NSTAGE) & 1));
      } else {
        const int nvalid = (B - b0) * AP;
        for (int i = tid; i < nvalid; i += NTHR) sbuf[i] = __ldg(&g_logits[(size_t)b0 * AP + i]);
        asm volatile("bar.sync 1, %0;" ::"r"(NTHR) : "memory");
      }
      const int b = b0 + slot;
      const bool row_ok = b < B;
      const long long r = (long long)(row_ok ? b : b0) * A + a;
      const float* lg = sbuf + (row_ok ? slot : 0) * AP + a * P;

      // ---- softmax terms (TF: max over the FINITE logits; a non-finite logit contributes nothing) -------------
      // Non-finite logits (and the unused 36th half-slot) are mapped to -FLT_MAX once: their term is then exactly 0 with
      // no further test.  e^d = 2^(d log2 e) straight on the MUFU pipe: |relative error| <= 2^-22 + |d| 6e-8, weighted by
      // the term itself that is < 5e-7 of the total -- well inside the budget `delta` of the margin test below.
      float e1[EPL];
      float m = -3.402823466e38f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const bool ok = i < 17 || c == 0;
        const float x = ok ? lg[kof(i)] : -3.402823466e38f;
        e1[i] = (fabsf(x) <= 3.402823466e38f && ok) ? x : kRoExcluded;
        m = fmaxf(m, e1[i]);
      }
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      const bool degenerate = !(m > kRoExcluded);  // no usable logit at all (or all below -1e30): the literal algorithm decides
      float T = 0.f, Hs = 0.f;  // block total (slots < 16); sum e (l - m) for the entropy
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const float d = e1[i] - m;  // (an excluded slot: -1e30 - m, finite, its term 2^(d log2 e) exactly 0 and 0 * d = -0)
        const float e = ex2f(d * kLog2e);
        Hs = fmaf(e, d, Hs);
        e1[i] = e;
        if (i < 16) T += e;  // running total inside this lane's block of 16 consecutive particles
      }
      // CDF in particle order: lane 0's block, lane 1's block, then particles 32, 33, 34
      const float To = __shfl_xor_sync(0xffffffffu, T, 1);
      const float base = c == 0 ? 0.f : To;
      const float T01 = T + To;
      const float o16 = __shfl_xor_sync(0xffffffffu, e1[16], 1), o17 = __shfl_xor_sync(0xffffffffu, e1[17], 1);
      const float e32 = c == 0 ? e1[16] : o16, e33 = c == 0 ? o16 : e1[16], e34 = c == 0 ? e1[17] : o17;
      const float c32 = T01 + e32, c33 = c32 + e33, c34 = c33 + e34;
      const float total = c34;
      Hs += __shfl_xor_sync(0xffffffffu, Hs, 1);
      // ---- the draw ---------------------------------------------------------------------------------------------
      // (lane 0 of the row generates the uniform's Philox block, lane 1 the normal's -- the same instructions, different
      //  counters -- and they swap through two shuffles: one block per lane instead of two)
      double u;
      uint32_t qnx = 0u, qny = 0u;
      if (kp.a.ext_uniform != nullptr) {
        u = kp.a.ext_uniform[r];
      } else {
        const uint4 q = rng(kp.a.offset + (uint64_t)c, (uint64_t)r);
        const uint32_t ox = __shfl_xor_sync(0xffffffffu, q.x, 1), oy = __shfl_xor_sync(0xffffffffu, q.y, 1);
        u = c == 0 ? u64_to_unit_double(q.x, q.y) : u64_to_unit_double(ox, oy);
        qnx = c == 0 ? ox : q.x;
        qny = c == 0 ? oy : q.y;
      }
      const float to_find = (float)(u * (double)total);
      const float margin = 2.f * delta * total + 1.2e-7f * total;
      // d_k = cdf_k - to_find, accumulated directly (run starts at base - to_find: one add per particle); the index is the
      // number of d_k < 0 -- the sign bits -- and the row is uncertain iff min |d_k| <= margin (d_k == 0 included, so the
      // difference between "< 0" and TF's "<= to_find" is always decided by the exact path).  Three instructions per particle.
      unsigned cnt = 0u;
      float dmin = 3.402823466e38f;
      float run = base - to_find;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        run += e1[i];
        cnt += __float_as_uint(run) >> 31;
        dmin = fminf(dmin, fabsf(run));
      }
      {  // the three tail values: lane 0 counts particles 32 and 34, lane 1 particle 33
        const float da = (c == 0 ? c32 : c33) - to_find;
        cnt += __float_as_uint(da) >> 31;
        dmin = fminf(dmin, fabsf(da));
        const float db = c == 0 ? c34 - to_find : 1.f;  // (lane 1: a positive dummy that cannot be the minimum... of interest)
        cnt += c == 0 ? __float_as_uint(db) >> 31 : 0u;
        dmin = c == 0 ? fminf(dmin, fabsf(db)) : dmin;
      }
      bool near = !(dmin > margin);  // (NaN -> near)
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
      {  // (no short-circuit around the shuffle: every lane must execute it)
        const int near_o = __shfl_xor_sync(0xffffffffu, (int)near, 1);
        near = near || near_o != 0;
      }
      int idx = (int)cnt;
      const bool fallback = near || idx >= P || !(total > 0.f) || !isfinite(total) || degenerate;
      // rows that need the literal fp64 algorithm (~0.1 %): one at a time, by the whole warp, in warp-uniform control flow
      unsigned fb_mask = __ballot_sync(0xffffffffu, fallback && c == 0);
      while (fb_mask != 0u) {
        const int src = __ffs(fb_mask) - 1;
        fb_mask &= fb_mask - 1;
        const int off = __shfl_sync(0xffffffffu, (int)(lg - sbuf), src);
        const unsigned ulo = __shfl_sync(0xffffffffu, (unsigned)__double2loint(u), src);
        const unsigned uhi = __shfl_sync(0xffffffffu, (unsigned)__double2hiint(u), src);
        const int k64 = ro_multinomial_fp64_warp(sbuf + off, P, __hiloint2double((int)uhi, (int)ulo), fb_s + warp * 40, lane);
        if ((lane & ~1) == src) idx = k64;  // both lanes of that row
      }
      __syncwarp();
      // ---- the action: Normal(loc, scale).sample()[idx] = eps * scale + loc (utils.py:190-194) -------------------
      float eps;
      if (kp.a.ext_normal != nullptr) {
        eps = __ldg(&kp.a.ext_normal[r * P + idx]);
      } else {
        eps = normal_from_bits(qnx, qny);  // (the function K2 uses: same bits)
      }
      const float2 ms_ = musd[a * P + idx];  // (shared-memory table: a dependent global load here was 5 % of all stall samples)
      const float v = __fadd_rn(__fmul_rn(eps, ms_.y), ms_.x);
      // ---- log_prob of the action (utils.py:108-134, plain variant), entropy (:146-151), statistics (a2c.py:346-365) ---
      const float is1 = rcpf(total);
      const float is1s = row_ok ? is1 : 0.f;  // (a masked row adds probability 0: max(vmax, 0) = vmax, vsum + 0)
      float S2 = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const float4 q4 = csp[i * TPS];
        const float z = fmaf(v, q4.x, q4.y);
        const float n = ex2f(fmaf(z * z, -0.5f * kLog2e, q4.z));
        S2 = fmaf(e1[i], n, S2);  // (e1 == 0 for "not a particle")
        const float pr = e1[i] * is1s;
        vmax[i] = fmaxf(vmax[i], pr);
        vsum[i] += pr;
      }
      __syncwarp();
      S2 += __shfl_xor_sync(0xffffffffu, S2, 1);
      const float l2t = lg2f(total);
      const float lnp = kLn2 * (lg2f(S2) - l2t);  // -inf when every term underflowed (p == 0)
      const float Hval = fmaf(kLn2, l2t, -Hs * is1);  // sum_k p_k (ln s1 - (l_k - m))
      if (c == 0) {
        rowbuf[st * SLOTS * A + slot * A + a] = row_ok ? make_float2(lnp, Hval) : make_float2(0.f, 0.f);
        if (row_ok) {
          kp.a.action[r] = v;
          kp.a.idx[r] = idx;
        }
      }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar[st])) : "memory");
    }
    // ---- per-thread statistics -> per-slot tables (in the stage ring: every compute thread is done reading it) -----
    asm volatile("bar.sync 1, %0;" ::"r"(NTHR) : "memory");
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      if (i < 17 || c == 0) {
        const int k = kof(i);
        red[(slot * 2 + 0) * AP + a * P + k] = vmax[i];
        red[(slot * 2 + 1) * AP + a * P + k] = vsum[i];
      }
    }
  }
  __syncthreads();
  if (kp.part != nullptr) {
    float* part = kp.part + (size_t)blockIdx.x * 2 * AP;
    for (int i = tid; i < AP; i += NTHR + 32) {
      float mx = 0.f, sm = 0.f;
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        mx = fmaxf(mx, red[(sl * 2 + 0) * AP + i]);
        sm += red[(sl * 2 + 1) * AP + i];
      }
      part[i] = mx;
      part[AP + i] = sm;
    }
  }
}

// max_active = max(max_active, max over CTAs), sum_active += sum over CTAs (fixed order)
__global__ void rollout_stats_finalize_kernel(const float* __restrict__ part, float* __restrict__ max_active,
                                              float* __restrict__ sum_active, int nparts, int AP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= AP) return;
  float mx = 0.f, sm = 0.f;
  for (int p0 = 0; p0 < nparts; p0 += 8) {
    float xm[8], xs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = p0 + u < nparts;
      xm[u] = ok ? __ldcg(&part[(size_t)(p0 + u) * 2 * AP + i]) : 0.f;
      xs[u] = ok ? __ldcg(&part[(size_t)(p0 + u) * 2 * AP + AP + i]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      mx = fmaxf(mx, xm[u]);
      sm += xs[u];
    }
  }
  max_active[i] = fmaxf(max_active[i], mx);
  sum_active[i] += sm;
}

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_rollout_workspace_bytes(int32_t A, int32_t P, size_t* bytes) {
  if (!bytes || A <= 0 || P <= 0) return PFPN_ERR_ARG;
  *bytes = (size_t)kRoMaxCtas * 2 * A * P * sizeof(float);
  return PFPN_OK;
}

extern "C" int pfpn_head_rollout(const pfpn_rollout_args* args, void* workspace, size_t workspace_bytes,
                                 pfpn_stream_t stream_) {
  if (!args) return PFPN_ERR_ARG;
  const pfpn_rollout_args& a = *args;
  if (a.B < 0 || a.A <= 0 || a.P <= 0) return PFPN_ERR_ARG;
  if (a.A != 36 || a.P != 35) return PFPN_ERR_UNSUPPORTED;  // the shipped DPPO-PFPN shape; others: the three-kernel form
  if (a.B == 0) return PFPN_OK;
  if (!a.logits || !a.loc || !a.logstd || !a.action || !a.idx || !a.lp) return PFPN_ERR_ARG;
  if ((a.max_active == nullptr) != (a.sum_active == nullptr)) return PFPN_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(a.logits) & 15u) return PFPN_ERR_ALIGN;
  const bool stats = a.max_active != nullptr;
  size_t need;
  pfpn_rollout_workspace_bytes(a.A, a.P, &need);
  if (stats && (!workspace || workspace_bytes < need)) return PFPN_ERR_WORKSPACE;
  constexpr int SLOTS = 4, NSTAGE = 3, AP = 36 * 35;  // (3 stages x 2 CTAs per SM = 120 KB of loads in flight per SM)
  constexpr int stage_bytes = (SLOTS * AP * 4 + 127) & ~127;
  constexpr int smem = NSTAGE * stage_bytes + 16 * NSTAGE + NSTAGE * SLOTS * 36 * 8 + 18 * 72 * 16 + AP * 8 + 9 * 40 * 8 + 128;
  int dev = 0, sms = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  PFPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RolloutK kp;
  kp.a = a;
  kp.part = stats ? reinterpret_cast<float*>(workspace) : nullptr;
  kp.num_tiles = (a.B + SLOTS - 1) / SLOTS;
  kp.wait_ns = pfpn_wait_ns(256u);
  int grid = 2 * sms;
  if (grid > kp.num_tiles) grid = kp.num_tiles;
  if (grid > kRoMaxCtas) grid = kRoMaxCtas;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  auto fn = rollout_kernel<SLOTS, NSTAGE>;
  PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  fn<<<grid, SLOTS * 72 + 32, smem, st>>>(kp);
  PFPN_CUDA_OK(cudaGetLastError());
  if (stats) {
    rollout_stats_finalize_kernel<<<(AP + 127) / 128, 128, 0, st>>>(kp.part, a.max_active, a.sum_active, grid, AP);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  return PFPN_OK;
}
