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

template <int LPR, int EPL, int RPT, int NSTAGE, bool BWD, int MAXT, int NREG>
__global__ void __launch_bounds__(MAXT) __maxnreg__(NREG) head_kernel(const HeadKParams kp) {
  constexpr int EP2 = (EPL + 1) / 2;  // packed pairs per lane
  constexpr int DIST = NSTAGE - 2;    // load prefetch distance (tiles)
  static_assert(NSTAGE >= 3, "need >=3 stages: 1 computing, >=1 loading, 1 draining");
  const pfpn_head_args& ar = kp.a;
  const int A = ar.A, P = ar.P, B = ar.B;
  const int AP = A * P;
  const int slots = kp.slots;
  const int TS = slots * RPT;
  const int tile_floats = TS * AP;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;

  // ---- shared memory carve-up ------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int stage_bytes = (tile_floats * 4 + 127) & ~127;
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  unsigned char* tail = smem_raw + (size_t)NSTAGE * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  float2* rowbuf = reinterpret_cast<float2*>(tail + 8 * NSTAGE);  // [2][TS*A]
  float* statebuf = reinterpret_cast<float*>(rowbuf + 2 * TS * A);  // [TS]
  float* lossbuf = statebuf + TS;                                   // [nwarps]

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
    mbar_fence_init();
  }

  // ---- fixed thread -> (slot, a, particle set) mapping ---------------------
  const int row_in_cta = tid / LPR;
  const int c = tid % LPR;
  const int slot = row_in_cta / A;
  const int a = row_in_cta - slot * A;
  const bool active = slot < slots;

  float2 isig[EP2], nmisig[EP2], cst[EP2];
  float2 acc1[EP2], acc2[EP2];
#pragma unroll
  for (int i2 = 0; i2 < EP2; ++i2) {
    float v[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = c + LPR * (2 * i2 + h);
      const bool ok = active && (2 * i2 + h < EPL) && (k < P);
      float ls = ok ? __ldg(&ar.logstd[a * P + k]) : 0.f;
      float mu = ok ? __ldg(&ar.loc[a * P + k]) : 0.f;
      float is = ok ? expf(-ls) : 0.f;
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
  const bool tanh_flag = (ar.flags & PFPN_HEAD_FLAG_TANH) != 0;
  const bool tail_exists = (B % TS) != 0;
  const int tail_tile = kp.num_tiles - 1;

  float adv_mean = 0.f, adv_rstd = 1.f;
  if (ar.mode == PFPN_HEAD_PPO && ar.adv_stats != nullptr) {
    adv_mean = __ldg(&ar.adv_stats[0]);
    adv_rstd = __ldg(&ar.adv_stats[1]);
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
    bulk_g2s(smem_u32(stage_base) + st * stage_bytes, ar.logits + (size_t)tile * tile_floats,
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
      dst[j] = (it < my_tiles && active && b < B) ? __ldg(&ar.value[(size_t)b * A + a]) : 0.f;
    }
  };
  load_values(0, v_nxt);

  int prev_tile = -1, prev_stage = 0;
  bool prev_was_tail = false;

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
        if (ar.mode == PFPN_HEAD_PPO) {
          pre_adv = __ldg(&ar.adv[b]);
          pre_lpo = __ldg(&ar.lp_old[b]);
        } else {
          pre_g = __ldg(&ar.g_lp[b]);
        }
      }
    }

    if (!is_tail) {
      mbar_wait(smem_u32(&full_bar[st]), (uint32_t)((it / NSTAGE) & 1));
    } else {
      const int nvalid = (B - b0) * AP;
      const float* src = ar.logits + (size_t)b0 * AP;
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
      const float* srow = sbuf + srow_idx * P + c;
      float2 l[EP2];
      float m = kNegBig;
#pragma unroll
      for (int i2 = 0; i2 < EP2; ++i2) {
        const int k0 = c + LPR * (2 * i2), k1 = k0 + LPR;
        l[i2].x = (row_ok && k0 < P) ? srow[LPR * (2 * i2)] : kNegBig;
        l[i2].y = (row_ok && (2 * i2 + 1 < EPL) && k1 < P) ? srow[LPR * (2 * i2 + 1)] : kNegBig;
        m = max3f(m, l[i2].x, l[i2].y);
      }
      m = row_max<LPR>(m);
      const float2 nmL = splat2(-m * kLog2e);
      const float2 L2 = splat2(kLog2e);
      const float2 nhl = splat2(-0.5f * kLog2e);
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
        if (ar.ent_ba != nullptr) ar.ent_ba[(size_t)(b0 + j * slots + slot) * A + a] = Hval;
      }
    }
    __syncthreads();  // B1: rowbuf complete; previous tile's gradient stores fenced

    // ---------------- thread 0: drain previous tile, prefetch ----------------
    if (tid == 0) {
      if (BWD && prev_tile >= 0) {
        bulk_s2g(ar.dlogits + (size_t)prev_tile * tile_floats,
                 smem_u32(stage_base) + prev_stage * stage_bytes, (uint32_t)(tile_floats * 4));
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
        ar.lp[b] = lp;
        if (ar.ent != nullptr) ar.ent[b] = en;
        if (BWD) {
          float g;
          if (ar.mode == PFPN_HEAD_PPO) {
            // (s == warp always holds while TS <= nwarps; otherwise re-read)
            const float adv_raw = (s == warp) ? pre_adv : __ldg(&ar.adv[b]);
            const float lpo = (s == warp) ? pre_lpo : __ldg(&ar.lp_old[b]);
            const float an = (adv_raw - adv_mean) * adv_rstd;
            const float ratio = expf(lp - lpo);
            const float surr = ratio * an;
            const float clipped = fminf(fmaxf(ratio, 1.f - ar.eps_clip), 1.f + ar.eps_clip) * an;
            loss_acc -= fminf(surr, clipped) * ar.loss_scale;
            g = (surr <= clipped) ? -ar.loss_scale * ratio * an : 0.f;  // TF Minimum: ties -> x
          } else {
            g = (s == warp) ? pre_g : __ldg(&ar.g_lp[b]);
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
        float* srow = sbuf + srow_idx * P + c;
        const float g = row_ok ? statebuf[sidx] : 0.f;
        const bool p_ok = s2r[j] > 0.f;
        // guard of utils.py:109-117: dL/dp is Inf/NaN when p == 0 -> zeroed
        const float g_row = p_ok ? g : 0.f;
        const float gs2 = p_ok ? g_row * rcpf(s2r[j]) : 0.f;
        const float2 v2 = splat2(v_cur[j]);
        const float2 gs2v = splat2(gs2);
        float dv = 0.f;
        if (!kp.has_ent_grad) {
          const float2 nc0 = splat2(-g_row * inv_s1[j]);
#pragma unroll
          for (int i2 = 0; i2 < EP2; ++i2) {
            const float2 rr = mul2(e2[j][i2], gs2v);  // g * r_k
            const float2 d = fma2(e1[j][i2], nc0, rr);
            const float2 z = fma2(v2, isig[i2], nmisig[i2]);
            const float2 q1 = fma2(z, z, splat2(-1.f));
            acc1[i2] = fma2(rr, z, acc1[i2]);
            acc2[i2] = fma2(rr, q1, acc2[i2]);
            if (ar.dvalue != nullptr) {
              const float2 w = mul2(mul2(rr, z), isig[i2]);
              dv += w.x + w.y;
            }
            const int k0 = c + LPR * (2 * i2), k1 = k0 + LPR;
            if (row_ok && k0 < P) srow[LPR * (2 * i2)] = d.x;
            if (row_ok && (2 * i2 + 1 < EPL) && k1 < P) srow[LPR * (2 * i2 + 1)] = d.y;
          }
        } else {
          float ge = ar.g_ent;
          if (ar.g_ent_ba != nullptr && row_ok) ge += __ldg(&ar.g_ent_ba[(size_t)b * A + a]);
          // dH/dl_k = -pi_k (ln pi_k + H),  ln pi_k = ln2*(t_k - log2 s1)
          const float c0 = (g_row + ge * (Hrow[j] - kLn2 * l2s1[j])) * inv_s1[j];
          const float2 nc0 = splat2(-c0);
          const float2 nc1 = splat2(-ge * kLn2 * inv_s1[j]);
#pragma unroll
          for (int i2 = 0; i2 < EP2; ++i2) {
            const float2 rr = mul2(e2[j][i2], gs2v);
            // t_k = log2(e1_k) would cost a MUFU; recover it from e1 is lossy, so
            // re-read the logit (still in smem) and recompute t = (l - m)*log2e via
            // t = log2(e1) only when e1 > 0 is not safe -> use lg2 of e1 guarded.
            float2 t;
            t.x = lg2f(fmaxf(e1[j][i2].x, 1e-37f));
            t.y = lg2f(fmaxf(e1[j][i2].y, 1e-37f));
            const float2 inner = fma2(t, nc1, nc0);
            const float2 d = fma2(e1[j][i2], inner, rr);
            const float2 z = fma2(v2, isig[i2], nmisig[i2]);
            const float2 q1 = fma2(z, z, splat2(-1.f));
            acc1[i2] = fma2(rr, z, acc1[i2]);
            acc2[i2] = fma2(rr, q1, acc2[i2]);
            if (ar.dvalue != nullptr) {
              const float2 w = mul2(mul2(rr, z), isig[i2]);
              dv += w.x + w.y;
            }
            const int k0 = c + LPR * (2 * i2), k1 = k0 + LPR;
            if (row_ok && k0 < P) srow[LPR * (2 * i2)] = d.x;
            if (row_ok && (2 * i2 + 1 < EPL) && k1 < P) srow[LPR * (2 * i2 + 1)] = d.y;
          }
        }
        if (ar.dvalue != nullptr) {
          dv = row_sum<LPR>(dv);
          if (row_ok && c == 0) {
            float out = -dv;
            if (tanh_flag) out += g_row * 2.f * tanhf(v_cur[j]);
            ar.dvalue[(size_t)b * A + a] = out;
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
        bulk_s2g(ar.dlogits + (size_t)prev_tile * tile_floats,
                 smem_u32(stage_base) + prev_stage * stage_bytes, (uint32_t)(tile_floats * 4));
        bulk_commit();
      }
    } else {
      const int b0 = prev_tile * TS;
      const int nvalid = (B - b0) * AP;
      const float* sb = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage_base) + (size_t)prev_stage * stage_bytes);
      float* dst = ar.dlogits + (size_t)b0 * AP;
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
        const int k0 = c + LPR * (2 * i2), k1 = k0 + LPR;
        if (k0 < P) {
          red[(slot * 2 + 0) * AP + a * P + k0] = acc1[i2].x;
          red[(slot * 2 + 1) * AP + a * P + k0] = acc2[i2].x;
        }
        if ((2 * i2 + 1 < EPL) && k1 < P) {
          red[(slot * 2 + 0) * AP + a * P + k1] = acc1[i2].y;
          red[(slot * 2 + 1) * AP + a * P + k1] = acc2[i2].y;
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
  int pmax, lpr, epl, rpt, nstage, maxt, nreg;
  head_kernel_t fwd, bwd;
};
#define PFPN_HEAD_VARIANT_ENTRY(PMAX, LPR, EPL, RPT, NST, MAXT, NREG)                                   \
  {                                                                                                     \
    PMAX, LPR, EPL, RPT, NST, MAXT, NREG, head_kernel<LPR, EPL, RPT, NST, false, MAXT, NREG>,           \
        head_kernel<LPR, EPL, RPT, NST, true, MAXT, NREG>                                               \
  }
static const HeadVariant kHeadVariants[] = {
    PFPN_HEAD_VARIANT_ENTRY(12, 4, 3, 2, 4, 320, 96),
    PFPN_HEAD_VARIANT_ENTRY(36, 4, 9, 2, 4, 288, 112),
    PFPN_HEAD_VARIANT_ENTRY(36, 4, 9, 1, 4, 288, 72),
    PFPN_HEAD_VARIANT_ENTRY(36, 4, 9, 1, 4, 288, 112),
    PFPN_HEAD_VARIANT_ENTRY(36, 4, 9, 2, 3, 288, 112),
    PFPN_HEAD_VARIANT_ENTRY(64, 8, 8, 2, 4, 320, 96),
    PFPN_HEAD_VARIANT_ENTRY(104, 8, 13, 1, 4, 288, 112),
    PFPN_HEAD_VARIANT_ENTRY(104, 8, 13, 1, 3, 288, 72),
    PFPN_HEAD_VARIANT_ENTRY(256, 16, 16, 1, 3, 384, 168),
};

static const HeadVariant* pick_variant(int P) {
  static const int want = []() {
    const char* e = getenv("PFPN_HEAD_VARIANT");
    return e ? atoi(e) : 0;
  }();
  const HeadVariant* first = nullptr;
  int seen = 0;
  for (const HeadVariant& v : kHeadVariants) {
    if (v.pmax < P) continue;
    if (first == nullptr) first = &v;
    if (v.pmax != first->pmax) break;
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
  L->smem_bytes = v.nstage * stage_bytes + 8 * v.nstage + 2 * L->ts * A * 8 + L->ts * 4 + kHeadMaxWarps * 4 + 16;
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
