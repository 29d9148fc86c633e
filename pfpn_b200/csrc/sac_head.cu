// K3f -- the SAC-PFPN head in ONE pass over logits[B, A, P]: reparameterised sample (Gumbel-softmax straight-through,
// forward), the tanh-squashed mixture log_prob of that sample (forward) and the backward of both, given dL/dsample
// (from the critics) and dL/dlog_prob.
//
// Reference: /root/reference/networks/utils.py:156-186 (sample: RelaxedOneHotCategorical, mask2 :164-171, mask :176-183)
// and :108-144 (log_prob on the tuple (sample, s_pre)); closed forms in SURVEY.md Appendix A1 / A2.  The boundary-faithful
// form is three launches (pfpn_head_rsample_fwd, pfpn_head_logprob with TANH + dvalue, pfpn_head_rsample_bwd) that read
// or write [B,A,P] five times and generate every Gumbel / normal draw twice; this kernel reads the logits once, writes the
// gradient once (8AP + 12A + 8 bytes per state) and draws once.  It is issue / MUFU bound, not HBM bound:
// per particle ~24 integer instructions of Philox, 8 MUFU operations (2 lg2 Gumbel, 2 Box-Muller, 2 tanh, 2 ex2).
//
// Structure (same skeleton as K1, csrc/head_logprob.cu, minus the cross-row dependency: everything here is row-local):
//  * persistent CTAs, one tile = SLOTS consecutive states fetched by ONE bulk async copy (TMA, UBLKCP) into a ring of
//    NSTAGE shared-memory stages; the gradient overwrites the logits in place and leaves with one bulk store issued
//    by the producer warp once every compute thread has arrived on the stage's `done` barrier.  No block-wide barrier in
//    the steady state; compute warps may drift NSTAGE-1 tiles apart.
//  * a mixture row (b, a) is owned by 8 adjacent lanes; lane c owns the particles k = 32 j + 4 u + 16 (c / 4) + c % 4
//    (u, j < 4): conflict-free at a 100-float row pitch (see K1) AND exactly the four particles one Philox4x32 block of
//    the split kernels serves (k0, k0+32, k0+64, k0+96 with k0 = 4 u + 16 (c/4) + c%4), so with the same (seed, offset)
//    this kernel draws the same Gumbel / normal variates as pfpn_head_rsample_fwd / _bwd.
//  * per-row reductions are 3-step xor butterflies over the 8 lanes: one (max, argmax, winner's p) and one of five sums.
//  * dloc / dlogstd: the mixture terms accumulate in registers (a thread's particle set never changes), the winner's
//    straight-through terms in a per-slot shared-memory table written by the row's own lanes only; per-CTA partials
//    are combined in CTA order by K1's finalize kernel -- no float atomics, bit-reproducible.
#include <string.h>

#include "common.cuh"

namespace pfpn {

constexpr uint64_t kSacNormalStream = 1ull << 63;  // same stream split as csrc/rsample.cu
constexpr int kSacMaxCtas = 148 * 2;

__device__ __forceinline__ float sac_sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sac_bits_to_unit(uint32_t w) {  // 23 random bits -> (i + 0.5) 2^-23, strictly in (0, 1)
  return __uint_as_float(0x3f800000u | (w >> 9)) - 0.99999994f;
}
__device__ __forceinline__ void sac_normal4(const uint4 q, float (&n)[4]) {  // two Box-Muller pairs, MUFU pipe
  const float r0 = sac_sqrt_approx(-2.f * kLn2 * lg2f(sac_bits_to_unit(q.x)));
  const float r1 = sac_sqrt_approx(-2.f * kLn2 * lg2f(sac_bits_to_unit(q.z)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * sac_bits_to_unit(q.y), &s0, &c0);
  __sincosf(6.283185307179586f * sac_bits_to_unit(q.w), &s1, &c1);
  n[0] = r0 * c0;
  n[1] = r0 * s0;
  n[2] = r1 * c1;
  n[3] = r1 * s1;
}
__device__ __forceinline__ float sac_fast_tanh(float x) {  // 1 - 2 / (1 + e^{2x}); |err| ~ 1e-7
  const float e = ex2f(x * (2.f * kLog2e));
  return 1.f - 2.f * rcpf(1.f + e);
}

constexpr int kSacRounds = 7;  // Philox4x32-7, as csrc/rsample.cu
struct SacHeadK {
  pfpn_sac_head_args a;
  float* part;  // [grid][2 * A * P]
  int num_tiles;
  uint32_t wait_ns;  // producer back-off while the compute threads work on a tile
  uint32_t rk[2 * kSacRounds];  // Philox round keys, precomputed on the host: they reach the XORs as constant-bank operands
};
// Philox4x32-7 with the key schedule taken from the kernel parameters (identical output to PhiloxR<7>(seed))
__device__ __forceinline__ uint4 sac_philox(const SacHeadK& kp, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < kSacRounds; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c0 = hi1 ^ c1 ^ kp.rk[2 * r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ kp.rk[2 * r + 1];
    c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// P = PT particles (32 < PT <= 128, so that every lane's four Philox blocks exist), A = AT action dims, 8 lanes per row.
template <int PT, int AT, int SLOTS, int NSTAGE, bool FAST>
__global__ void __launch_bounds__(SLOTS* AT * 8 + 32, 1) sac_head_kernel(const SacHeadK kp) {
  constexpr int P = PT, A = AT, AP = A * P, LPR = 8;
  constexpr int NE = 13;                       // particle slots per lane: (u, j<3) -> 3u + j, (u = 0, j = 3) -> 12
  constexpr int NP = (NE + 1) / 2;             // ... held as register pairs for the packed fp32x2 instructions
  constexpr int NTHR = SLOTS * A * LPR;        // compute threads
  constexpr int TILE_F = SLOTS * AP;           // logits floats per tile
  constexpr int STAGE_BYTES = ((TILE_F + SLOTS * A) * 4 + 127) & ~127;
  static_assert(PT > 96 && PT <= 100 + 0 * AT, "instantiated for P = 100 (slot (0,3) valid for c < P - 96 <= 4)");
  static_assert((A * LPR) % 32 == 0, "rows must not straddle warps");
  const int B = kp.a.B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == NTHR / 32;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* tail = smem_raw + (size_t)NSTAGE * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* done_bar = full_bar + NSTAGE;
  float* rowbuf = reinterpret_cast<float*>(tail + 16 * NSTAGE);  // [NSTAGE][SLOTS * A] per-row log p
  float2* ms_s = reinterpret_cast<float2*>(rowbuf + NSTAGE * SLOTS * A);  // [AP] {mu, sigma}
  float2* mi_s = ms_s + AP;                                                 // [AP] {-mu, 1 / sigma}
  float* cst_s = reinterpret_cast<float*>(mi_s + AP);                       // [AP] -(logstd + ln sqrt(2 pi)) log2 e
  float* acc_s = cst_s + AP;  // [SLOTS][2][AP] straight-through terms of the winning particles

  if (is_producer && lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&done_bar[s]), (uint32_t)NTHR);
    }
    mbar_fence_init();
  }
  // The two float2 tables are stored PERMUTED inside every 32-particle block: an 8-byte access is served per half warp
  // (2 rows x 8 lanes), rows are 100 float2 = 4 (mod 16) slots apart, so the 8 lanes of a row must cover {0..3, 8..11}
  // (+ 4 for odd u): pos(32 j + 4 u + 16 h + w) = 32 j + 16 (u / 2) + 4 (u % 2) + 8 h + w.  (61 M bank conflicts per
  // launch in the first round-2 profile came from these two tables in natural order.)
  auto perm = [](int k) -> int {
    const int r = k & 31;
    return (k & ~31) + 16 * ((r >> 3) & 1) + 4 * ((r >> 2) & 1) + 8 * (r >> 4) + (r & 3);
  };
  for (int i = tid; i < AP; i += NTHR + 32) {
    const float ls = __ldg(&kp.a.logstd[i]), mu = __ldg(&kp.a.loc[i]);
    const int ai = i / P, ki = i - ai * P;
    const int ip = ai * P + perm(ki);
    ms_s[ip] = make_float2(mu, FAST ? ex2f(ls * kLog2e) : expf(ls));  // (FAST: the scale the split kernels use -> the same locations)
    mi_s[ip] = make_float2(-mu, expf(-ls));
    cst_s[i] = -(ls + kHalfLog2Pi) * kLog2e;
  }
  for (int i = tid; i < SLOTS * 2 * AP; i += NTHR + 32) acc_s[i] = 0.f;
  __syncthreads();

  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  int my_tiles = 0;
  if (first_tile < kp.num_tiles) my_tiles = (kp.num_tiles - 1 - first_tile) / tile_step + 1;
  const bool tail_exists = (B % SLOTS) != 0;
  const int tail_tile = kp.num_tiles - 1;
  const float* __restrict__ g_logits = kp.a.logits;
  float* __restrict__ g_dlogits = kp.a.dlogits;
  unsigned char* stage0 = smem_raw;

  if (is_producer) {
    // ===================== TMA producer warp =====================================================================
    auto issue_load = [&](int it) {
      const int tile = first_tile + it * tile_step;
      if (it >= my_tiles || (tail_exists && tile == tail_tile)) return;  // the ragged last tile is copied cooperatively
      const int st = it % NSTAGE;
      const uint32_t bar = smem_u32(&full_bar[st]);
      mbar_expect_tx(bar, (uint32_t)((TILE_F + SLOTS * A) * 4));
      bulk_g2s(smem_u32(stage0) + st * STAGE_BYTES, g_logits + (size_t)tile * TILE_F, (uint32_t)(TILE_F * 4), bar);
      bulk_g2s(smem_u32(stage0) + st * STAGE_BYTES + TILE_F * 4, kp.a.g_sample + (size_t)tile * SLOTS * A,
               (uint32_t)(SLOTS * A * 4), bar);
    };
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < NSTAGE; ++d) issue_load(d);
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % NSTAGE;
      const int tile = first_tile + it * tile_step;
      mbar_wait_sleep(smem_u32(&done_bar[st]), (uint32_t)((it / NSTAGE) & 1), kp.wait_ns);  // every compute thread finished (and fenced) this tile
      if (lane < SLOTS) {  // log_prob of the state = sum over a of the per-row log p, fixed order
        const int b = tile * SLOTS + lane;
        if (b < B) {
          const float* rb = rowbuf + st * SLOTS * A + lane * A;
          float lp = 0.f;
#pragma unroll 4
          for (int aa = 0; aa < A; ++aa) lp += rb[aa];
          kp.a.logp[b] = lp;
        }
      }
      __syncwarp();
      if (lane == 0) {
        if (!(tail_exists && tile == tail_tile)) {
          bulk_s2g(g_dlogits + (size_t)tile * TILE_F, smem_u32(stage0) + st * STAGE_BYTES, (uint32_t)(TILE_F * 4));
          bulk_commit();
          bulk_wait_read<0>();  // the stage may be refilled once the store has READ it
        }
        issue_load(it + NSTAGE);
      }
      __syncwarp();
    }
  } else {
    // ===================== compute threads ========================================================================
    const int slot = tid / (A * LPR);
    const int rem = tid - slot * (A * LPR);
    const int a = rem >> 3, c = rem & 7;
    const int kb0 = 16 * (c >> 2) + (c & 3);
    const int pb0 = 8 * (c >> 2) + (c & 3);  // this lane's offset inside a permuted 32-block of the float2 tables
    const bool tail_ok = c < P - 96;  // slot (u = 0, j = 3): particle 96 + kb0
    // (opaque copies: under the 96-register cap the compiler otherwise re-derives slot / a / c / kb0 / pb0 from
    //  threadIdx.x at every use inside the tile loop -- 5 % of the executed instructions in the first profile)
    int kb0_p = kb0, pb0_p = pb0, rowoff_p = slot * AP + a * P, sa_p = slot * A + a;
    asm volatile("" : "+r"(kb0_p), "+r"(pb0_p), "+r"(rowoff_p), "+r"(sa_p));
    const float2* ms_r = ms_s + a * P;
    const float2* mi_r = mi_s + a * P;
    const float* cst_r = cst_s + a * P;
    float* acc_r = acc_s + (size_t)slot * 2 * AP + a * P;
    float2 acc1[NP], acc2[NP];
    const uint64_t rng_off = kp.a.offset + (kp.a.offset_dev != nullptr ? *kp.a.offset_dev : 0ull);  // (device word: graph replays)
#pragma unroll
    for (int i = 0; i < NP; ++i) acc1[i] = acc2[i] = make_float2(0.f, 0.f);

    int st = 0;
    uint32_t ph = 0u;
    for (int it = 0, tile = first_tile; it < my_tiles; ++it, tile += tile_step, ph ^= (st == NSTAGE - 1), st = (st == NSTAGE - 1) ? 0 : st + 1) {
      float* sbuf = reinterpret_cast<float*>(stage0 + (size_t)st * STAGE_BYTES);
      const bool is_tail = tail_exists && tile == tail_tile;
      const int b0 = tile * SLOTS;
      if (!is_tail) {
        mbar_wait(smem_u32(&full_bar[st]), ph);
      } else {
        // (the slots past the batch end are zero-filled: their threads run the row math on their OWN slot -- finite inputs,
        //  results discarded -- instead of re-reading slot 0's row, which slot 0's threads overwrite with the gradient in the
        //  same phase: a benign but real read-after-write race compute-sanitizer's racecheck reported)
        const int nvalid = (B - b0) * AP;
        for (int i = tid; i < TILE_F; i += NTHR) sbuf[i] = i < nvalid ? __ldg(&g_logits[(size_t)b0 * AP + i]) : 0.f;
        for (int i = tid; i < SLOTS * A; i += NTHR)
          sbuf[TILE_F + i] = i < (B - b0) * A ? __ldg(&kp.a.g_sample[(size_t)b0 * A + i]) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"r"(NTHR) : "memory");  // compute threads only; the last tile only
      }
      const int b = b0 + slot;
      const bool row_ok = b < B;
      const long long r = (long long)(row_ok ? b : b0) * A + a;  // (masked rows: zero logits, draws of state b0, results discarded)
      float* lg = sbuf + rowoff_p;

      // ---- step 1: draws, noisy logits (log2 domain), locations -----------------------------------------------
      // Slots are kept as register PAIRS (slot e -> pair e / 2, half e % 2) so that steps 3 / 4 run on packed fp32x2
      // instructions; the 14th half-slot is a permanently masked dummy.
      float2 y2[NP], nl2[NP], p2[NP], e22[NP];
      y2[NP - 1].y = -3.402823466e38f;
      nl2[NP - 1].y = 0.f;
      p2[NP - 1].y = 0.f;
      float ymax = -3.402823466e38f, yfm = 0.f, pmax = 0.f;
      int kmax = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kb = 4 * u + kb0_p;
        const int nj = u == 0 ? 4 : 3;  // particles >= 100: only slot (0, 3) exists
        float n4[4], nl4[4], ya4[4];
        if (FAST) {
          const uint4 q = sac_philox(kp, rng_off, (uint64_t)(r * P + kb));
          sac_normal4(sac_philox(kp, rng_off, kSacNormalStream | (uint64_t)(r * P + kb)), n4);
          const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nj) nl4[j] = fmaxf(-lg2f(sac_bits_to_unit(w4[j])), 4e-8f);  // -log2 u, kept off 0
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j >= nj) continue;
            const int k = kb + 32 * j;
            const bool ok = j < 3 || tail_ok;
            const float U = ok ? __ldg(&kp.a.ext_uniform[r * P + k]) : 0.5f;
            n4[j] = ok ? __ldg(&kp.a.ext_normal[r * P + k]) : 0.f;
            nl4[j] = fmaxf(-log2f(U), 4e-8f);
            ya4[j] = -logf(-logf(U));  // the reference's Gumbel, natural-log domain: decides the argmax bit-exactly
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j >= nj) continue;
          const int e = j < 3 ? 3 * u + j : 12;
          const int k = kb + 32 * j;
          const bool ok = j < 3 || tail_ok;
          const float x = ok ? lg[k] : 0.f;
          const float2 ms = ms_r[ok ? 32 * j + 16 * (u >> 1) + 4 * (u & 1) + pb0_p : 0];  // {mu, sigma}, permuted table
          const float yf = ok ? fmaf(x, kLog2e, -(FAST ? lg2f(nl4[j]) : log2f(nl4[j]))) : -3.402823466e38f;
          const float ycmp = FAST ? yf : (ok ? x + ya4[j] : -3.402823466e38f);  // (G + logits) / T, T = 1
          const float pk = FAST ? fmaf(n4[j], ms.y, ms.x) : __fadd_rn(__fmul_rn(n4[j], ms.y), ms.x);
          if (e & 1) {
            nl2[e >> 1].y = nl4[j];
            y2[e >> 1].y = yf;
            p2[e >> 1].y = pk;
          } else {
            nl2[e >> 1].x = nl4[j];
            y2[e >> 1].x = yf;
            p2[e >> 1].x = pk;
          }
          const bool take = FAST ? (ycmp > ymax) : (ycmp > ymax || (ycmp == ymax && k < kmax));  // ties -> smallest index (tf.argmax)
          if (take) {
            ymax = ycmp;
            if (!FAST) yfm = yf;  // (FAST: the comparison key IS yf)
            kmax = k;
            pmax = pk;
          }
        }
      }
      // ---- step 2: row argmax over the 8 lanes, carrying the winner's location --------------------------------
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float oy = __shfl_xor_sync(0xffffffffu, ymax, o), of = FAST ? 0.f : __shfl_xor_sync(0xffffffffu, yfm, o);
        const float op = __shfl_xor_sync(0xffffffffu, pmax, o);
        const int ok_ = __shfl_xor_sync(0xffffffffu, kmax, o);
        const bool take = oy > ymax || (oy == ymax && ok_ < kmax);
        if (take) {
          ymax = oy;
          if (!FAST) yfm = of;
          kmax = ok_;
          pmax = op;
        }
      }
      const float uu = pmax;  // s_pre: the winner's pre-tanh location
      const int arg = kmax;
      // tanh and 1 - tanh^2 without cancellation: e = exp(-2|u|), sech^2 = 4 e / (1 + e)^2
      const float e2m = ex2f(-2.f * kLog2e * fabsf(uu)), q1 = rcpf(1.f + e2m);
      const float omt2 = 4.f * e2m * q1 * q1;
      const float t = copysignf(1.f - 2.f * e2m * q1, uu);

      // ---- step 3: particle terms and the five row sums (packed fp32x2) ------------------------------------------
      // slot e of this lane <-> particle k = 4 u + kb0 + 32 j, (u, j) = (e / 3, e % 3) for e < 12, (0, 3) for e = 12
      auto slot_k = [&](int e) -> int {
        const int u = e < 12 ? e / 3 : 0, j = e < 12 ? e % 3 : 3;
        const int k = 4 * u + kb0_p + 32 * j;
        return (e < 12 || (e == 12 && tail_ok)) ? k : 0;  // masked half-slots read entry 0 (their terms are exactly 0)
      };
      auto slot_p = [&](int e) -> int {  // position of slot e's particle in the permuted float2 tables
        const int u = e < 12 ? e / 3 : 0, j = e < 12 ? e % 3 : 3;
        const int q = 32 * j + 16 * (u >> 1) + 4 * (u & 1) + pb0_p;
        return (e < 12 || (e == 12 && tail_ok)) ? q : 0;
      };
      if (FAST) yfm = ymax;
      const float2 uu2 = splat2(uu), nyf2 = splat2(-yfm), one2 = splat2(1.f), mtwo2 = splat2(-2.f);
      const float2 tl2 = splat2(2.f * kLog2e), nhl2 = splat2(-0.5f * kLog2e);
      float2 S1v = splat2(0.f), S2v = S1v, Tv = S1v, Swv = S1v, Swtv = S1v;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int k0 = slot_k(2 * i), k1 = slot_k(2 * i + 1);
        const float2 m0 = mi_r[slot_p(2 * i)], m1 = mi_r[slot_p(2 * i + 1)];  // {-mu, 1 / sigma}
        const float2 nmu = make_float2(m0.x, m1.x), is2 = make_float2(m0.y, m1.y);
        const float2 cs2 = make_float2(cst_r[k0], cst_r[k1]);
        const float2 dy = add2(y2[i], nyf2);
        const float2 w = make_float2(ex2f(dy.x), ex2f(dy.y));  // softmax(logits + G) numerator; 0 for a masked slot
        const float2 e1 = mul2(w, nl2[i]);                     // = 2^(log2e logit - yfm): softmax(logits) numerator, no 2nd ex2
        const float2 a2 = mul2(p2[i], tl2);
        const float2 den = add2(make_float2(ex2f(a2.x), ex2f(a2.y)), one2);
        const float2 th = fma2(make_float2(rcpf(den.x), rcpf(den.y)), mtwo2, one2);  // tanh p = 1 - 2 / (1 + e^{2p})
        const float2 z = mul2(add2(uu2, nmu), is2);
        const float2 t2 = fma2(mul2(z, z), nhl2, cs2);
        const float2 x2 = mul2(e1, make_float2(ex2f(t2.x), ex2f(t2.y)));
        y2[i] = w;
        p2[i] = th;
        e22[i] = x2;
        S1v = add2(S1v, e1);
        S2v = add2(S2v, x2);
        Tv = fma2(mul2(x2, z), is2, Tv);
        Swv = add2(Swv, w);
        Swtv = fma2(w, th, Swtv);
      }
      float S1 = S1v.x + S1v.y, S2 = S2v.x + S2v.y, T = Tv.x + Tv.y, Sw = Swv.x + Swv.y, Swt = Swtv.x + Swtv.y;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        S1 += __shfl_xor_sync(0xffffffffu, S1, o);
        S2 += __shfl_xor_sync(0xffffffffu, S2, o);
        T += __shfl_xor_sync(0xffffffffu, T, o);
        Sw += __shfl_xor_sync(0xffffffffu, Sw, o);
        Swt += __shfl_xor_sync(0xffffffffu, Swt, o);
      }
      // ---- step 4: row scalars, gradients ---------------------------------------------------------------------------
      const float is1 = rcpf(S1);
      // ln(1 - tanh^2 u) = ln 4 - 2|u| - 2 ln(1 + e^{-2|u|})   (utils.py:132-133 in a cancellation-free form)
      const float corr = 2.f * kLn2 - 2.f * fabsf(uu) - 2.f * kLn2 * lg2f(1.f + e2m);
      const float lnp = kLn2 * (lg2f(S2) - lg2f(S1)) - corr;  // -inf when every term underflowed (p == 0)
      const bool p_ok = S2 > 0.f;
      const float g = row_ok ? __ldg(&kp.a.g_lp[b]) : 0.f;
      const float g_row = p_ok ? g : 0.f;                       // guard of utils.py:109-117
      const float gs2 = p_ok ? g_row * rcpf(S2) : 0.f;
      const float ga = row_ok ? sbuf[TILE_F + sa_p] : 0.f;
      const float gu = fmaf(g_row, 2.f * t, -T * gs2);          // dL/du through log_prob: g (-sum r z / sigma + 2 tanh u)
      const float coef = fmaf(gu, rcpf(fmaxf(1e-6f, omt2)), ga);  // mask + mask2 (utils.py:164-183)
      const float iw = rcpf(Sw);
      const float wd = coef * fmaf(Swt, iw, -t);                // sum_j w_j D_j,  D_j = (tanh p_j - t) coef
      const float2 gs2v = splat2(gs2), iwv = splat2(iw), coefv = splat2(coef), nt2 = splat2(-t), nwd2 = splat2(-wd);
      const float2 nc1 = splat2(-g_row * is1), neg1 = splat2(-1.f);
      float* stp = row_ok ? lg : nullptr;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int k0 = slot_k(2 * i), k1 = slot_k(2 * i + 1);
        const float2 m0 = mi_r[slot_p(2 * i)], m1 = mi_r[slot_p(2 * i + 1)];
        const float2 z = mul2(add2(uu2, make_float2(m0.x, m1.x)), make_float2(m0.y, m1.y));
        const float2 w = y2[i];
        const float2 rr = mul2(e22[i], gs2v);                                           // g r_k
        const float2 X = fma2(coefv, add2(p2[i], nt2), nwd2);                          // D_k - sum_j w_j D_j
        const float2 d = fma2(mul2(w, iwv), X, fma2(mul2(w, nl2[i]), nc1, rr));        // g (r - pi) + w (D - sum w D)
        acc1[i] = fma2(rr, z, acc1[i]);
        acc2[i] = fma2(rr, fma2(z, z, neg1), acc2[i]);
        if (stp != nullptr) {
          if (2 * i < 12 || tail_ok) stp[k0] = d.x;  // (slot 12 exists on lanes c < 4 only)
          if (2 * i + 1 < 12) stp[k1] = d.y;
        }
      }
      if (c == 0 && row_ok) {
        // the winner's straight-through terms: dL/dp_{k*} = (1 - t^2) g_a + g_u; this row's lanes are the only writers of
        // acc_r (slot, a) -- plain read-modify-write.  Stored times sigma: the finalize kernel divides dloc by sigma.
        const int ra = arg & 31;
        const int parg = (arg & ~31) + 16 * ((ra >> 3) & 1) + 4 * ((ra >> 2) & 1) + 8 * (ra >> 4) + (ra & 3);
        const float gp = fmaf(omt2, ga, gu) * ms_r[parg].y;
        const float zs = (uu + mi_r[parg].x) * mi_r[parg].y;      // = eps of the winner
        acc_r[arg] += gp;
        acc_r[AP + arg] += gp * zs;
        const size_t o = (size_t)b * A + a;
        kp.a.sample[o] = t;
        kp.a.s_pre[o] = uu;
        kp.a.idx[o] = arg;
      }
      if (c == 0) rowbuf[st * SLOTS * A + sa_p] = row_ok ? lnp : 0.f;
      if (is_tail) {  // ragged last tile: plain stores of the valid part
        asm volatile("bar.sync 1, %0;" ::"r"(NTHR) : "memory");
        const int nvalid = (B - b0) * AP;
        for (int i = tid; i < nvalid; i += NTHR) g_dlogits[(size_t)b0 * AP + i] = sbuf[i];
      }
      fence_async_smem();  // gradient STS -> visible to the bulk store
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar[st])) : "memory");
    }
    // ---- fold the register accumulators into this slot's table (own entries only) --------------------------------
    __syncwarp();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int u = e < 12 ? e / 3 : 0, j = e < 12 ? e % 3 : 3;
      const int k = 4 * u + kb0 + 32 * j;
      if (e < 12 || tail_ok) {
        acc_r[k] += (e & 1) ? acc1[e >> 1].y : acc1[e >> 1].x;
        acc_r[AP + k] += (e & 1) ? acc2[e >> 1].y : acc2[e >> 1].x;
      }
    }
  }
  __syncthreads();  // (the producer has waited for every bulk store's read; all tables complete)
  float* part = kp.part + (size_t)blockIdx.x * 2 * AP;
  for (int i = tid; i < 2 * AP; i += NTHR + 32) {
    float s = 0.f;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) s += acc_s[(size_t)sl * 2 * AP + i];
    part[i] = s;
  }
}

}  // namespace pfpn

using namespace pfpn;

// (defined in head_logprob.cu) deterministic second stage over per-CTA [2*AP] partials: dloc = sum / sigma, dlogstd = sum
extern "C" int pfpn_head_finalize_partials(const float* part, int32_t nparts, const float* logstd, float* dloc, float* dlogstd,
                                           int32_t AP, pfpn_stream_t stream);

namespace {
constexpr int kSacSlots = 2, kSacStages = 3;
template <bool FAST>
int sac_launch(const SacHeadK& kp, int grid, cudaStream_t st) {
  constexpr int P = 100, A = 36, AP = A * P;
  constexpr int stage_bytes = ((kSacSlots * AP + kSacSlots * A) * 4 + 127) & ~127;
  constexpr int smem = kSacStages * stage_bytes + 16 * kSacStages + kSacStages * kSacSlots * A * 4 + 5 * AP * 4 +
                       kSacSlots * 2 * AP * 4 + 128;
  auto fn = sac_head_kernel<P, A, kSacSlots, kSacStages, FAST>;
  PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  fn<<<grid, kSacSlots * A * 8 + 32, smem, st>>>(kp);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}
}  // namespace

extern "C" int pfpn_sac_head_workspace_bytes(int32_t A, int32_t P, size_t* bytes) {
  if (!bytes || A <= 0 || P <= 0) return PFPN_ERR_ARG;
  *bytes = (size_t)kSacMaxCtas * 2 * A * P * sizeof(float);
  return PFPN_OK;
}

extern "C" int pfpn_sac_head_fwd_bwd(const pfpn_sac_head_args* args, void* workspace, size_t workspace_bytes,
                                     pfpn_stream_t stream_) {
  if (!args) return PFPN_ERR_ARG;
  const pfpn_sac_head_args& a = *args;
  if (a.B < 0 || a.A <= 0 || a.P <= 0) return PFPN_ERR_ARG;
  if (a.A != 36 || a.P != 100) return PFPN_ERR_UNSUPPORTED;  // the BASELINE c5 shape; other shapes: the three-launch form
  if (a.B == 0) return PFPN_OK;
  if (!a.logits || !a.loc || !a.logstd || !a.g_sample || !a.g_lp || !a.sample || !a.s_pre || !a.idx || !a.logp || !a.dlogits ||
      !a.dloc || !a.dlogstd)
    return PFPN_ERR_ARG;
  if ((a.ext_uniform == nullptr) != (a.ext_normal == nullptr)) return PFPN_ERR_ARG;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (!al16(a.logits) || !al16(a.dlogits) || !al16(a.g_sample) || !al16(workspace)) return PFPN_ERR_ALIGN;
  size_t need;
  pfpn_sac_head_workspace_bytes(a.A, a.P, &need);
  if (!workspace || workspace_bytes < need) return PFPN_ERR_WORKSPACE;
  int dev = 0, sms = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  PFPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SacHeadK kp;
  kp.a = a;
  kp.part = reinterpret_cast<float*>(workspace);
  kp.num_tiles = (a.B + kSacSlots - 1) / kSacSlots;
  kp.wait_ns = pfpn_wait_ns(256u);
  {
    uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    for (int r = 0; r < kSacRounds; ++r) {
      kp.rk[2 * r] = k0;
      kp.rk[2 * r + 1] = k1;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
  }
  int grid = sms;  // one 19-warp CTA per SM
  if (grid > kp.num_tiles) grid = kp.num_tiles;
  if (grid > kSacMaxCtas) grid = kSacMaxCtas;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  int rc = a.ext_uniform == nullptr ? sac_launch<true>(kp, grid, st) : sac_launch<false>(kp, grid, st);
  if (rc != PFPN_OK) return rc;
  return pfpn_head_finalize_partials(kp.part, grid, a.logstd, a.dloc, a.dlogstd, a.A * a.P, stream_);
}
