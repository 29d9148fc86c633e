// K1 -- fused PFPN head: mixture log_prob + categorical entropy (+ PPO clipped
// surrogate) forward AND backward in one pass over logits[B, A, P].
//
// Math follows SURVEY.md Appendix A1, i.e. the closed form of what
// /root/reference/networks/utils.py:108-151 + ppo.py:44-54 build as ~70 TF
// graph nodes.  This file shares no structure with the reference (which has no
// kernels at all); it is a from-scratch sm_100a design:
//
//  * persistent CTAs; each iteration handles one TILE of TS consecutive states,
//    i.e. one contiguous TS*A*P*4-byte chunk of `logits`;
//  * the chunk is fetched with ONE 1-D bulk async copy (TMA, SASS UBLKCP) into
//    a ring of NSTAGE shared-memory buffers tracked by mbarriers, gradients are
//    written over the logits in place and leave with ONE bulk store, so HBM
//    sees exactly 4AP bytes in + 4AP bytes out per state, fully coalesced;
//  * a mixture row (b,a) is owned by LPR adjacent lanes, EPL interleaved
//    particles each; the per-row max / sum reductions are LPR-wide xor-shuffle
//    butterflies; per-(a,k) constants and the dloc/dlogstd accumulators stay in
//    registers for the whole kernel because a thread's (a, k-set) never changes;
//  * element math is packed fp32x2 (FFMA2/FMUL2/FADD2) with 2 MUFU.EX2 per
//    particle; everything is computed in the log2 domain;
//  * sum over a (log_prob of the state) and the PPO dL/dlp go through a tiny
//    smem exchange: 2 block barriers per tile.
#include <stdlib.h>

#include "common.cuh"

namespace pfpn {

constexpr int kHeadMaxWarps = 32;
constexpr float kNegBig = -1.0e30f;

struct HeadKParams {
  pfpn_head_args a;
  float* part;       // [grid][2*A*P]  per-CTA dloc/dlogstd partial sums
  float* loss_part;  // [grid]
  int num_tiles;
  int slots;         // state slots per CTA; threads = slots*A*LPR (rounded to 32)
  int has_ent_grad;
};

template <int LPR>
__device__ __forceinline__ float row_max(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float softplusf_acc(float x) {
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

template <int LPR, int EPL, int RPT, int NSTAGE, bool BWD, int MAXT, int NREG, int PT>
__global__ void __launch_bounds__(MAXT) __maxnreg__(NREG) head_kernel(const HeadKParams kp) {
  constexpr int EP2 = (EPL + 1) / 2;  // packed pairs per lane
  constexpr int DIST = NSTAGE - 2;    // load prefetch distance (tiles)
  static_assert(NSTAGE >= 3, "need >=3 stages: 1 computing, >=1 loading, 1 draining");
  // hot parameters into registers / uniform registers once
  const int A = kp.a.A, B = kp.a.B;
  const int P = PT > 0 ? PT : kp.a.P;  // PT > 0: particle count known at compile time
  const int AP = A * P;
  const int slots = kp.slots;
  const int TS = slots * RPT;
  const int tile_floats = TS * AP;
  const uint32_t mode = kp.a.mode;
  const float* __restrict__ g_logits = kp.a.logits;
  const float* __restrict__ g_value = kp.a.value;
  float* __restrict__ g_dlogits = kp.a.dlogits;
  float* __restrict__ g_dvalue = kp.a.dvalue;
  float* __restrict__ g_ent_ba = kp.a.ent_ba;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;

  // ---- shared memory carve-up ------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int stage_bytes = (tile_floats * 4 + 127) & ~127;
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  unsigned char* tail = smem_raw + (size_t)NSTAGE * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  float2* rowbuf = reinterpret_cast<float2*>(tail + 8 * NSTAGE);    // [2][TS*A]
  float* statebuf = reinterpret_cast<float*>(rowbuf + 2 * TS * A);  // [TS]
  float* lossbuf = statebuf + TS;                                   // [kHeadMaxWarps]
  float* dummy = lossbuf + kHeadMaxWarps;                           // [LPR*EPL] sink for masked rows

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
    mbar_fence_init();
  }

  // ---- fixed thread -> (slot, a, particle set) mapping ---------------------
  // lane c of a row owns particles k = c + LPR*i, i = 0..EPL-1.  Slots i < nfull are
  // valid for every lane, slot nfull only for c < P % LPR, later slots for nobody.
  const int row_in_cta = tid / LPR;
  const int c = tid % LPR;
  const int slot = row_in_cta / A;
  const int a = row_in_cta - slot * A;
  const bool active = slot < slots;
  const int nfull = P / LPR;
  const bool part_ok = c < P - nfull * LPR;
  auto k_ok = [&](int i) -> bool { return (i < nfull) || (i == nfull && part_ok); };

  float2 isig[EP2], nmisig[EP2], cst[EP2];
  float2 acc1[EP2], acc2[EP2];
#pragma unroll
  for (int i2 = 0; i2 < EP2; ++i2) {
    float v[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * i2 + h;
      const bool ok = active && (i < EPL) && k_ok(i);
      const int k = c + LPR * i;
      const float ls = ok ? __ldg(&kp.a.logstd[a * P + k]) : 0.f;
      const float mu = ok ? __ldg(&kp.a.loc[a * P + k]) : 0.f;
      const float is = ok ? expf(-ls) : 0.f;
      v[h][0] = is;
      v[h][1] = -mu * is;
      v[h][2] = ok ? -(ls + kHalfLog2Pi) * kLog2e : 0.f;
    }
    isig[i2] = make_float2(v[0][0], v[1][0]);
    nmisig[i2] = make_float2(v[0][1], v[1][1]);
    cst[i2] = make_float2(v[0][2], v[1][2]);
    acc1[i2] = make_float2(0.f, 0.f);
    acc2[i2] = make_float2(0.f, 0.f);
  }
  const bool tanh_flag = (kp.a.flags & PFPN_HEAD_FLAG_TANH) != 0;
  const bool tail_exists = (B % TS) != 0;
  const int tail_tile = kp.num_tiles - 1;
  const bool has_ent_grad = kp.has_ent_grad != 0;
  const float eps_clip = kp.a.eps_clip, loss_scale = kp.a.loss_scale;

  float adv_mean = 0.f, adv_rstd = 1.f;
  if (mode == PFPN_HEAD_PPO && kp.a.adv_stats != nullptr) {
    adv_mean = __ldg(&kp.a.adv_stats[0]);
    adv_rstd = __ldg(&kp.a.adv_stats[1]);
  }
  float loss_acc = 0.f;

  __syncthreads();  // mbarrier init visible

  const int first_tile = blockIdx.x;
  const int tile_step = gridDim.x;
  int my_tiles = 0;
  if (first_tile < kp.num_tiles) my_tiles = (kp.num_tiles - 1 - first_tile) / tile_step + 1;

  auto issue_load = [&](int it) {  // thread 0 only
    const int tile = first_tile + it * tile_step;
    if (it >= my_tiles) return;
    if (tail_exists && tile == tail_tile) return;  // tail is copied cooperatively
    const int st = it % NSTAGE;
    const uint32_t bar = smem_u32(&full_bar[st]);
    mbar_expect_tx(bar, (uint32_t)(tile_floats * 4));
    bulk_g2s(smem_u32(stage_base) + st * stage_bytes, g_logits + (size_t)tile * tile_floats,
             (uint32_t)(tile_floats * 4), bar);
  };
  if (tid == 0) {
#pragma unroll
    for (int d = 0; d < DIST; ++d) issue_load(d);
  }

  // value prefetch (one tile ahead)
  float v_nxt[RPT];
  auto load_values = [&](int it, float (&dst)[RPT]) {
    const int tile = first_tile + it * tile_step;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int b = tile * TS + j * slots + slot;
      dst[j] = (it < my_tiles && active && b < B) ? __ldg(&g_value[(size_t)b * A + a]) : 0.f;
    }
  };
  load_values(0, v_nxt);

  int prev_tile = -1, prev_stage = 0;
  bool prev_was_tail = false;

  const float2 L2 = splat2(kLog2e);
  const float2 nhl = splat2(-0.5f * kLog2e);
  const float2 neg1 = splat2(-1.f);

  for (int it = 0; it < my_tiles; ++it) {
    const int tile = first_tile + it * tile_step;
    const int st = it % NSTAGE;
    float* sbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage_base) + (size_t)st * stage_bytes);
    const bool is_tail = tail_exists && tile == tail_tile;
    const int b0 = tile * TS;

    float v_cur[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) v_cur[j] = v_nxt[j];
    load_values(it + 1, v_nxt);

    // per-state scalars for the reducer lanes, fetched early to hide latency
    float pre_g = 0.f, pre_adv = 0.f, pre_lpo = 0.f;
    if (BWD && lane == 0 && warp < TS) {
      const int b = b0 + warp;
      if (b < B) {
        if (mode == PFPN_HEAD_PPO) {
          pre_adv = __ldg(&kp.a.adv[b]);
          pre_lpo = __ldg(&kp.a.lp_old[b]);
        } else {
          pre_g = __ldg(&kp.a.g_lp[b]);
        }
      }
    }

    if (!is_tail) {
      mbar_wait(smem_u32(&full_bar[st]), (uint32_t)((it / NSTAGE) & 1));
    } else {
      const int nvalid = (B - b0) * AP;
      const float* src = g_logits + (size_t)b0 * AP;
      for (int idx = tid; idx < nvalid; idx += nthr) sbuf[idx] = __ldg(&src[idx]);
      __syncthreads();
    }

    // ---------------- pass A/B: per-row statistics ------------------------
    float2 e1[RPT][EP2], e2[RPT][EP2];
    float inv_s1[RPT], s2r[RPT], Hrow[RPT], l2s1[RPT];
    float2* rb = rowbuf + (it & 1) * TS * A;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int srow_idx = (j * slots + slot) * A + a;
      const bool row_ok = active && (b0 + j * slots + slot < B);
      // masked rows read state 0 of the tile (always valid, finite) and are discarded
      const float* ld = sbuf + (row_ok ? srow_idx : a) * P + c;
      float2 l[EP2];
      float m = kNegBig;
#pragma unroll
      for (int i2 = 0; i2 < EP2; ++i2) {
        const int i0 = 2 * i2, i1 = 2 * i2 + 1;
        l[i2].x = (i0 < nfull) ? ld[LPR * i0] : ((i0 == nfull && part_ok) ? ld[LPR * i0] : kNegBig);
        l[i2].y = (i1 >= EPL) ? kNegBig
                              : ((i1 < nfull) ? ld[LPR * i1] : ((i1 == nfull && part_ok) ? ld[LPR * i1] : kNegBig));
        m = max3f(m, l[i2].x, l[i2].y);
      }
      m = row_max<LPR>(m);
      const float2 nmL = splat2(-m * kLog2e);
      const float2 v2 = splat2(v_cur[j]);
      float2 s1 = make_float2(0.f, 0.f), s2 = s1, h = s1;
#pragma unroll
      for (int i2 = 0; i2 < EP2; ++i2) {
        const float2 t = fma2(l[i2], L2, nmL);
        float2 x1;
        x1.x = ex2f(t.x);
        x1.y = ex2f(t.y);
        const float2 z = fma2(v2, isig[i2], nmisig[i2]);
        const float2 q = mul2(z, z);
        const float2 u = add2(t, cst[i2]);
        const float2 t2 = fma2(q, nhl, u);
        float2 x2;
        x2.x = ex2f(t2.x);
        x2.y = ex2f(t2.y);
        s1 = add2(s1, x1);
        s2 = add2(s2, x2);
        h = fma2(x1, t, h);
        e1[j][i2] = x1;
        e2[j][i2] = x2;
      }
      const float S1 = row_sum<LPR>(s1.x + s1.y);
      const float S2 = row_sum<LPR>(s2.x + s2.y);
      const float Hs = row_sum<LPR>(h.x + h.y);
      const float is1 = rcpf(S1);
      const float lg1 = lg2f(S1);
      const float Hval = kLn2 * (lg1 - Hs * is1);
      float lnp = kLn2 * (lg2f(S2) - lg1);  // -inf when every term underflowed (p == 0)
      if (tanh_flag) {
        const float uu = v_cur[j];
        lnp -= 2.f * (kLn2 - uu - softplusf_acc(-2.f * uu));
      }
      inv_s1[j] = is1;
      s2r[j] = S2;
      Hrow[j] = Hval;
      l2s1[j] = lg1;
      if (row_ok && c == 0) {
        rb[srow_idx] = make_float2(lnp, Hval);
        if (g_ent_ba != nullptr) g_ent_ba[(size_t)(b0 + j * slots + slot) * A + a] = Hval;
      }
    }
    __syncthreads();  // B1: rowbuf complete; previous tile's gradient stores fenced

    // ---------------- thread 0: drain previous tile, prefetch ----------------
    if (tid == 0) {
      if (BWD && prev_tile >= 0) {
        bulk_s2g(g_dlogits + (size_t)prev_tile * tile_floats, smem_u32(stage_base) + prev_stage * stage_bytes,
                 (uint32_t)(tile_floats * 4));
        bulk_commit();
        bulk_wait_read<1>();
      }
      issue_load(it + DIST);
    }

    // ---------------- per-state reduction over a (+ PPO) ---------------------
    for (int s = warp; s < TS; s += nwarps) {
      const int b = b0 + s;
      float lp = 0.f, en = 0.f;
      for (int aa = lane; aa < A; aa += 32) {
        const float2 r = rb[s * A + aa];
        lp += r.x;
        en += r.y;
      }
      lp = row_sum<32>(lp);
      en = row_sum<32>(en);
      if (lane == 0 && b < B) {
        kp.a.lp[b] = lp;
        if (kp.a.ent != nullptr) kp.a.ent[b] = en;
        if (BWD) {
          float g;
          if (mode == PFPN_HEAD_PPO) {
            // (s == warp always holds while TS <= nwarps; otherwise re-read)
            const float adv_raw = (s == warp) ? pre_adv : __ldg(&kp.a.adv[b]);
            const float lpo = (s == warp) ? pre_lpo : __ldg(&kp.a.lp_old[b]);
            const float an = (adv_raw - adv_mean) * adv_rstd;
            const float ratio = expf(lp - lpo);
            const float surr = ratio * an;
            const float clipped = fminf(fmaxf(ratio, 1.f - eps_clip), 1.f + eps_clip) * an;
            loss_acc -= fminf(surr, clipped) * loss_scale;
            g = (surr <= clipped) ? -loss_scale * ratio * an : 0.f;  // TF Minimum: ties -> x
          } else {
            g = (s == warp) ? pre_g : __ldg(&kp.a.g_lp[b]);
          }
          statebuf[s] = g;
        }
      }
    }

    if (BWD) {
      __syncthreads();  // B2: statebuf ready
      // ---------------- pass C: gradients ----------------------------------
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int sidx = j * slots + slot;
        const int srow_idx = sidx * A + a;
        const int b = b0 + sidx;
        const bool row_ok = active && (b < B);
        float* stp = row_ok ? (sbuf + srow_idx * P + c) : (dummy + c);
        const float g = row_ok ? statebuf[sidx] : 0.f;
        const bool p_ok = s2r[j] > 0.f;
        // guard of utils.py:109-117: dL/dp is Inf/NaN when p == 0 -> zeroed
        const float g_row = p_ok ? g : 0.f;
        const float gs2 = p_ok ? g_row * rcpf(s2r[j]) : 0.f;
        const float2 v2 = splat2(v_cur[j]);
        const float2 gs2v = splat2(gs2);
        float2 nc0, nc1;
        if (!has_ent_grad) {
          nc0 = splat2(-g_row * inv_s1[j]);
          nc1 = splat2(0.f);
        } else {
          float ge = kp.a.g_ent;
          if (kp.a.g_ent_ba != nullptr && row_ok) ge += __ldg(&kp.a.g_ent_ba[(size_t)b * A + a]);
          // dH/dl_k = -pi_k (ln pi_k + H),  ln pi_k = ln2*(t_k - log2 s1)
          nc0 = splat2(-(g_row + ge * (Hrow[j] - kLn2 * l2s1[j])) * inv_s1[j]);
          nc1 = splat2(-ge * kLn2 * inv_s1[j]);
        }
#pragma unroll
        for (int i2 = 0; i2 < EP2; ++i2) {
          const float2 rr = mul2(e2[j][i2], gs2v);  // g * r_k
          float2 coef = nc0;
          if (has_ent_grad) {
            // t_k = (l_k - m) log2e recovered as log2(e1_k); e1 == 0 contributes nothing
            float2 t;
            t.x = lg2f(fmaxf(e1[j][i2].x, 1e-37f));
            t.y = lg2f(fmaxf(e1[j][i2].y, 1e-37f));
            coef = fma2(t, nc1, nc0);
          }
          const float2 d = fma2(e1[j][i2], coef, rr);
          const float2 z = fma2(v2, isig[i2], nmisig[i2]);
          const float2 q1 = fma2(z, z, neg1);
          acc1[i2] = fma2(rr, z, acc1[i2]);
          acc2[i2] = fma2(rr, q1, acc2[i2]);
          const int i0 = 2 * i2, i1 = 2 * i2 + 1;
          if (i0 < nfull) stp[LPR * i0] = d.x;
          else if (i0 == nfull && part_ok) stp[LPR * i0] = d.x;
          if (i1 < EPL) {
            if (i1 < nfull) stp[LPR * i1] = d.y;
            else if (i1 == nfull && part_ok) stp[LPR * i1] = d.y;
          }
        }
        if (g_dvalue != nullptr) {
          float dv = 0.f;
#pragma unroll
          for (int i2 = 0; i2 < EP2; ++i2) {
            const float2 rr = mul2(e2[j][i2], gs2v);
            const float2 z = fma2(v2, isig[i2], nmisig[i2]);
            const float2 w = mul2(mul2(rr, z), isig[i2]);
            dv += w.x + w.y;
          }
          dv = row_sum<LPR>(dv);
          if (row_ok && c == 0) {
            float out = -dv;
            if (tanh_flag) out += g_row * 2.f * tanhf(v_cur[j]);
            g_dvalue[(size_t)b * A + a] = out;
          }
        }
      }
      fence_async_smem();  // make this thread's gradient STS visible to the bulk store
      prev_tile = tile;
      prev_stage = st;
      prev_was_tail = is_tail;
    }
  }

  // ---------------- epilogue: last store, partial sums --------------------------
  __syncthreads();
  if (BWD && prev_tile >= 0) {
    if (!prev_was_tail) {
      if (tid == 0) {
        bulk_s2g(g_dlogits + (size_t)prev_tile * tile_floats, smem_u32(stage_base) + prev_stage * stage_bytes,
                 (uint32_t)(tile_floats * 4));
        bulk_commit();
      }
    } else {
      const int b0 = prev_tile * TS;
      const int nvalid = (B - b0) * AP;
      const float* sb = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage_base) + (size_t)prev_stage * stage_bytes);
      float* dst = g_dlogits + (size_t)b0 * AP;
      for (int idx = tid; idx < nvalid; idx += nthr) dst[idx] = sb[idx];
    }
  }
  if (tid == 0) bulk_wait_read<0>();
  if (BWD) {
    if (lane == 0) lossbuf[warp] = loss_acc;
    __syncthreads();  // all bulk reads of smem done (thread 0 waited) before reuse
    float* red = stage_base;  // [slots][2][AP]
    if (active) {
#pragma unroll
      for (int i2 = 0; i2 < EP2; ++i2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * i2 + h;
          if (i < EPL && k_ok(i)) {
            const int k = c + LPR * i;
            red[(slot * 2 + 0) * AP + a * P + k] = h ? acc1[i2].y : acc1[i2].x;
            red[(slot * 2 + 1) * AP + a * P + k] = h ? acc2[i2].y : acc2[i2].x;
          }
        }
      }
    }
    __syncthreads();
    float* part = kp.part + (size_t)blockIdx.x * 2 * AP;
    for (int idx = tid; idx < 2 * AP; idx += nthr) {
      float s = 0.f;
      for (int sl = 0; sl < slots; ++sl) s += red[sl * 2 * AP + idx];
      part[idx] = s;
    }
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += lossbuf[w];
      kp.loss_part[blockIdx.x] = s;
    }
  }
}

// Deterministic second stage: sums the per-CTA partials in CTA order.
__global__ void head_finalize_kernel(const float* __restrict__ part, const float* __restrict__ loss_part,
                                     const float* __restrict__ logstd, float* __restrict__ dloc,
                                     float* __restrict__ dlogstd, float* __restrict__ loss, int AP, int nparts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < AP) {
    float s1 = 0.f, s2 = 0.f;
    for (int p = 0; p < nparts; ++p) {
      s1 += part[(size_t)p * 2 * AP + idx];
      s2 += part[(size_t)p * 2 * AP + AP + idx];
    }
    dloc[idx] = s1 * expf(-logstd[idx]);
    dlogstd[idx] = s2;
  }
  if (loss != nullptr && idx == 0) {
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += loss_part[p];
    *loss = s;
  }
}

// One CTA: mean and 1/(sqrt(popvar)+1e-8) of adv[B], fixed summation order.
__global__ void adv_stats_kernel(const float* __restrict__ adv, int B, float* __restrict__ stats) {
  __shared__ double sh[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0, ss = 0.0;
  for (int i = tid; i < B; i += blockDim.x) {
    const double x = adv[i];
    s += x;
    ss += x * x;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) {
    sh[0][warp] = s;
    sh[1][warp] = ss;
  }
  __syncthreads();
  if (tid == 0) {
    double S = 0.0, SS = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      S += sh[0][w];
      SS += sh[1][w];
    }
    const double mean = S / B;
    double var = SS / B - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[0] = (float)mean;
    stats[1] = (float)(1.0 / (sqrt(var) + 1e-8));
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef void (*head_kernel_t)(const HeadKParams);

// Compiled instantiations.  `pmax` = largest P the (LPR, EPL) pair covers; the
// first entry whose pmax >= P is the default, PFPN_HEAD_VARIANT=<n> picks the
// n-th matching entry instead (tuning aid, read once per process).
struct HeadVariant {
  int pmin, pmax, lpr, epl, rpt, nstage, maxt, nreg, pt;
  head_kernel_t fwd, bwd;
};
#define PFPN_HEAD_VARIANT_ENTRY(PMIN, PMAX, LPR, EPL, RPT, NST, MAXT, NREG, PT)                            \
  {                                                                                                        \
    PMIN, PMAX, LPR, EPL, RPT, NST, MAXT, NREG, PT, head_kernel<LPR, EPL, RPT, NST, false, MAXT, NREG, PT>, \
        head_kernel<LPR, EPL, RPT, NST, true, MAXT, NREG, PT>                                              \
  }
// [pmin, pmax] = particle counts an entry serves; PT > 0 entries are specialised for
// exactly P == PT (the shapes BASELINE.json names), PT == 0 entries take P at run time.
static const HeadVariant kHeadVariants[] = {
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 4, 9, 1, 4, 288, 72, 35),
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 4, 9, 1, 6, 288, 72, 35),
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 4, 9, 2, 4, 288, 96, 35),
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 4, 9, 1, 4, 288, 96, 35),
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 4, 9, 2, 3, 288, 72, 35),
    PFPN_HEAD_VARIANT_ENTRY(100, 100, 8, 13, 1, 4, 288, 96, 100),
    PFPN_HEAD_VARIANT_ENTRY(100, 100, 8, 13, 1, 3, 288, 72, 100),
    PFPN_HEAD_VARIANT_ENTRY(10, 10, 4, 3, 2, 4, 320, 72, 10),
    PFPN_HEAD_VARIANT_ENTRY(1, 12, 4, 3, 2, 4, 320, 96, 0),
    PFPN_HEAD_VARIANT_ENTRY(13, 36, 4, 9, 1, 4, 288, 96, 0),
    PFPN_HEAD_VARIANT_ENTRY(37, 64, 8, 8, 1, 4, 320, 96, 0),
    PFPN_HEAD_VARIANT_ENTRY(65, 104, 8, 13, 1, 4, 288, 96, 0),
    PFPN_HEAD_VARIANT_ENTRY(105, 256, 16, 16, 1, 3, 384, 168, 0),
};

static const HeadVariant* pick_variant(int P) {
  static const int want = []() {
    const char* e = getenv("PFPN_HEAD_VARIANT");
    return e ? atoi(e) : 0;
  }();
  const HeadVariant* first = nullptr;
  int seen = 0;
  for (const HeadVariant& v : kHeadVariants) {
    if (P < v.pmin || P > v.pmax) continue;
    if (first == nullptr) first = &v;
    if (v.pt != first->pt) break;
    if (seen == want) return &v;
    ++seen;
  }
  return first;
}

struct HeadLaunch {
  const HeadVariant* cfg;
  head_kernel_t fn;
  int threads, slots, ts, smem_bytes, ctas_per_sm, num_sms;
};

static int plan_head(int A, int P, bool bwd, HeadLaunch* L) {
  if (A <= 0 || P <= 0) return PFPN_ERR_ARG;
  L->cfg = pick_variant(P);
  if (L->cfg == nullptr) return PFPN_ERR_UNSUPPORTED;
  const HeadVariant& v = *L->cfg;
  const int per_slot = A * v.lpr;
  if (per_slot > v.maxt) return PFPN_ERR_UNSUPPORTED;
  int slots = v.maxt / per_slot;
  // tiles must start 16-byte aligned: TS*A*P % 4 == 0
  while (slots > 0 && ((slots * v.rpt * A * P) & 3) != 0) --slots;
  if (slots < 1) return PFPN_ERR_UNSUPPORTED;
  L->slots = slots;
  L->ts = slots * v.rpt;
  L->threads = (slots * per_slot + 31) & ~31;
  const int stage_bytes = (L->ts * A * P * 4 + 127) & ~127;
  L->smem_bytes = v.nstage * stage_bytes + 8 * v.nstage + 2 * L->ts * A * 8 + L->ts * 4 + kHeadMaxWarps * 4 +
                  v.lpr * v.epl * 4 + 16;
  L->fn = bwd ? v.bwd : v.fwd;
  int dev = 0;
  PFPN_CUDA_OK(cudaGetDevice(&dev));
  int max_optin = 0;
  PFPN_CUDA_OK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (L->smem_bytes > max_optin) return PFPN_ERR_UNSUPPORTED;
  PFPN_CUDA_OK(cudaFuncSetAttribute((const void*)L->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, L->smem_bytes));
  PFPN_CUDA_OK(cudaDeviceGetAttribute(&L->num_sms, cudaDevAttrMultiProcessorCount, dev));
  PFPN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L->ctas_per_sm, (const void*)L->fn, L->threads,
                                                             L->smem_bytes));
  if (L->ctas_per_sm < 1) return PFPN_ERR_UNSUPPORTED;
  return PFPN_OK;
}

constexpr int kMaxPartCtas = 148 * 8 + 64;

}  // namespace pfpn

using namespace pfpn;

extern "C" int pfpn_head_workspace_bytes(int32_t A, int32_t P, size_t* bytes) {
  if (bytes == nullptr || A <= 0 || P <= 0) return PFPN_ERR_ARG;
  *bytes = (size_t)kMaxPartCtas * (2 * (size_t)A * P + 1) * sizeof(float);
  return PFPN_OK;
}

extern "C" int pfpn_head_launch_info(int32_t A, int32_t P, uint32_t mode, int32_t* out) {
  if (out == nullptr) return PFPN_ERR_ARG;
  HeadLaunch L;
  int rc = plan_head(A, P, mode != PFPN_HEAD_FWD, &L);
  if (rc != PFPN_OK) return rc;
  out[0] = L.num_sms;
  out[1] = L.ctas_per_sm;
  out[2] = L.threads;
  out[3] = L.ts;
  return PFPN_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int pfpn_head_logprob(const pfpn_head_args* args, void* workspace, size_t workspace_bytes,
                                 pfpn_stream_t stream_) {
  if (args == nullptr) return PFPN_ERR_ARG;
  const pfpn_head_args& a = *args;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a.B < 0 || a.A <= 0 || a.P <= 0) return PFPN_ERR_ARG;
  if (a.mode > PFPN_HEAD_PPO) return PFPN_ERR_ARG;
  if (a.B == 0) return PFPN_OK;
  if (!a.logits || !a.loc || !a.logstd || !a.value || !a.lp) return PFPN_ERR_ARG;
  const bool bwd = a.mode != PFPN_HEAD_FWD;
  if (bwd && (!a.dlogits || !a.dloc || !a.dlogstd)) return PFPN_ERR_ARG;
  if (a.mode == PFPN_HEAD_GRAD && !a.g_lp) return PFPN_ERR_ARG;
  if (a.mode == PFPN_HEAD_PPO && (!a.adv || !a.lp_old || !a.loss)) return PFPN_ERR_ARG;
  if (!aligned16(a.logits) || (bwd && !aligned16(a.dlogits))) return PFPN_ERR_ALIGN;

  HeadLaunch L;
  int rc = plan_head(a.A, a.P, bwd, &L);
  if (rc != PFPN_OK) return rc;
  const int num_tiles = (a.B + L.ts - 1) / L.ts;
  int grid = L.num_sms * L.ctas_per_sm;
  if (grid > num_tiles) grid = num_tiles;
  if (grid > kMaxPartCtas) grid = kMaxPartCtas;

  HeadKParams kp;
  kp.a = a;
  kp.num_tiles = num_tiles;
  kp.slots = L.slots;
  kp.has_ent_grad = (a.g_ent != 0.f || a.g_ent_ba != nullptr) ? 1 : 0;
  kp.part = nullptr;
  kp.loss_part = nullptr;
  const size_t AP = (size_t)a.A * a.P;
  if (bwd) {
    const size_t need = (size_t)grid * (2 * AP + 1) * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) return PFPN_ERR_WORKSPACE;
    kp.part = reinterpret_cast<float*>(workspace);
    kp.loss_part = kp.part + (size_t)grid * 2 * AP;
  }
  L.fn<<<grid, L.threads, L.smem_bytes, stream>>>(kp);
  PFPN_CUDA_OK(cudaGetLastError());
  if (bwd) {
    const int thr = 256;
    const int blocks = (int)((AP + thr - 1) / thr);
    head_finalize_kernel<<<blocks, thr, 0, stream>>>(kp.part, kp.loss_part, a.logstd, a.dloc, a.dlogstd,
                                                     a.mode == PFPN_HEAD_PPO ? a.loss : nullptr, (int)AP, grid);
    PFPN_CUDA_OK(cudaGetLastError());
  }
  return PFPN_OK;
}

extern "C" int pfpn_adv_stats(const float* adv, int32_t B, float* stats, pfpn_stream_t stream_) {
  if (!adv || !stats || B <= 0) return PFPN_ERR_ARG;
  adv_stats_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(adv, B, stats);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}
