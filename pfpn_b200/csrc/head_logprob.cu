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
//  * sum over a (log_prob of the state) goes through 4-row group partials in shared memory; the
//    otherwise idle TMA producer warp sums them per state, writes lp / ent, evaluates the PPO
//    surrogate and publishes ONE dL/dlp per state through an mbarrier (SIP) -- there is no
//    block-wide barrier in the steady state.
#include <stdlib.h>
#include <string.h>

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

// KM (kernel mode): 0 = forward only; 1 = backward, every feature decided at run time;
// 2 / 3 = lean PPO / GRAD backward (no tanh, no entropy gradient, no dvalue, no ent_ba:
// the DPPO train step) with those branches compiled out.
// OPT bit 0 (CSM): per-(a,k) constants live in shared memory instead of registers.
// OPT bit 1 (RC):  the backward pass does not carry the particle terms e1_k, e2_k of a tile in registers
//                  across the pipeline step; pass C re-reads the (still intact) logits from the stage and
//                  recomputes them (+4 MUFU.EX2 per pair on a 30 %-busy pipe).  That frees the registers a
//                  2-lanes-per-row mapping needs (18 particles per lane), which halves the per-row fixed work.
template <int LPR, int EPL, int RPT, int NSTAGE, int KM, int MAXT, int NREG, int PT, int AT, int OPT>
__global__ void __launch_bounds__(((MAXT + 31) / 32) * 32 + 32) __maxnreg__(NREG) head_kernel(const HeadKParams kp) {
  constexpr bool BWD = KM != 0;
  constexpr bool CSM = (OPT & 1) != 0;
  constexpr bool RC = (OPT & 2) != 0 && BWD;
  // OPT bit 2 (SPL): bank-conflict-free particle ownership for 2 lanes per row.  Rows are P floats apart (P odd), so the
  // interleaved ownership k = c + 2 i makes lane pairs of different rows collide on a shared-memory bank (12.9 M
  // conflicts per launch at P = 35 in the round-1 profile).  With k = 16 c + i for i < 16 the two lanes of a row are 16
  // banks apart and the 32 lanes of a warp (16 consecutive rows, 3 banks apart each) hit 32 distinct banks; particles
  // >= 32 stay interleaved (P = 35).  8 lanes per row, rows 100 floats (4 banks) apart: k = 32 (i / 4) + 4 (i % 4) +
  // 16 (c / 4) + c % 4 -- the four rows of a warp land on banks 4r + {0..3, 16..19}, all 32 distinct in every round
  // (P = 100).  Both maps keep the validity rule "slot i < P / LPR for every lane, slot P / LPR for c < P % LPR".
  constexpr bool SPL = (OPT & 4) != 0;
  // OPT bit 3 (HINT): barrier waits carry a suspend-time hint -- 17 % of the warp instructions of the round-2 profile were
  // SYNCS / BRA / YIELD of spinning try_wait loops
  constexpr bool HINT = (OPT & 8) != 0;
  static_assert(!SPL || (LPR == 2 && PT == 35) || (LPR == 8 && PT == 100), "split ownership: P=35 / 2 lanes or P=100 / 8 lanes");
  constexpr bool LEAN = KM >= 2;
  constexpr int EP2 = (EPL + 1) / 2;  // packed pairs per lane
  // Software pipeline: iteration `it` runs pass A/B of tile it and pass C of tile it-1; the
  // gradient tile it-2 leaves with a bulk store.  A stage is therefore busy for 3 iterations
  // after its load completes, plus DIST iterations of prefetch.
  constexpr int DIST = NSTAGE - 3;
  static_assert(NSTAGE >= 4, "need >= 4 stages: load-ahead, A/B, C, draining store");
  // hot parameters into registers / uniform registers once
  // PT/AT > 0: particle count / action dims known at compile time (the BASELINE shapes);
  // every index expression below then folds to constants.
  const int B = kp.a.B;
  const int A = AT > 0 ? AT : kp.a.A;
  const int P = PT > 0 ? PT : kp.a.P;
  const int AP = A * P;
  const int slots = AT > 0 ? MAXT / (AT * LPR) : kp.slots;
  // SEG: a state's rows end on a 16-lane boundary, so the sum over a can be done with
  // half-warp partials (4 rows each) + one 16-lane butterfly instead of A/LPR strided reads.
  constexpr int GL = 4 * LPR;  // lanes of a 4-row group (a state ends on a group boundary when A % 4 == 0)
  constexpr bool SEG = AT > 0 && (AT % 4 == 0) && GL <= 32;
  constexpr int HPS = SEG ? AT / 4 : 1;  // 4-row groups per state
  // SIP (scalars in the producer): the otherwise idle producer warp sums a state's row partials, writes lp / ent, evaluates
  // the PPO surrogate and publishes ONE dL/dlp per state; the compute lanes (72..576 per state) no longer each re-derive it.
  constexpr bool SIP = SEG;
  const int TS = slots * RPT;
  const int tile_floats = TS * AP;
  const uint32_t mode = KM == 2 ? (uint32_t)PFPN_HEAD_PPO : (KM == 3 ? (uint32_t)PFPN_HEAD_GRAD : kp.a.mode);
  const float* __restrict__ g_logits = kp.a.logits;
  const float* __restrict__ g_value = kp.a.value;
  float* __restrict__ g_dlogits = kp.a.dlogits;
  float* __restrict__ g_dvalue = LEAN ? nullptr : kp.a.dvalue;
  float* __restrict__ g_ent_ba = LEAN ? nullptr : kp.a.ent_ba;
  const int tid = threadIdx.x;
  // the last warp of the CTA is the TMA producer (loads, gradient stores, their waits); the
  // first nthr threads compute
  const int nthr = blockDim.x - 32;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const bool is_producer = warp == nwarps;

  // ---- shared memory carve-up ------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // a stage = the logits tile followed by the tile's action values [TS][A] (fetched by the same producer
  // with a second bulk copy when A % 4 == 0, i.e. 16-byte granular; otherwise read with LDG one tile ahead)
  constexpr bool VAL_TMA = AT > 0 && (AT % 4 == 0);
  const int stage_bytes = (tile_floats * 4 + TS * A * 4 + 127) & ~127;
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  unsigned char* tail = smem_raw + (size_t)NSTAGE * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* cta_bar = full_bar + NSTAGE;                            // split-phase CTA barrier
  uint64_t* g_bar = cta_bar + 1;                                    // [2] SIP: producer -> compute, per-state dL/dlp ready (tile parity)
  float2* rowbuf = reinterpret_cast<float2*>(tail + 8 * (NSTAGE + 4));  // [NSTAGE][TS*A] per-row (log p, H) partials
  float* lossbuf = reinterpret_cast<float*>(rowbuf + NSTAGE * TS * A);  // [kHeadMaxWarps]
  float* gbuf = lossbuf + kHeadMaxWarps;                            // [NSTAGE][TS] per-state dL/dlp (SIP), 256 floats
  float* dummy = gbuf + 256;                                        // [LPR*EPL] sink for masked rows
  float2* cs = reinterpret_cast<float2*>(dummy + LPR * EPL + ((LPR * EPL) & 1));  // CSM: [EP2*3][A*LPR]

  if (is_producer && lane == 0) {  // the thread that also issues the first loads below, before anyone else is ready
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
    mbar_init(smem_u32(cta_bar), (uint32_t)nthr);
    mbar_init(smem_u32(&g_bar[0]), 1);
    mbar_init(smem_u32(&g_bar[1]), 1);
    mbar_fence_init();
  }
  // programmatic dependent launch: everything above overlaps the tail of the previous kernel on the stream
  // (adv_stats in the PPO step); no global memory is touched before this point
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const bool tail_exists = (B % TS) != 0;
  const int tail_tile = kp.num_tiles - 1;
  const int first_tile = blockIdx.x;
  const int tile_step = gridDim.x;
  int my_tiles = 0;
  if (first_tile < kp.num_tiles) my_tiles = (kp.num_tiles - 1 - first_tile) / tile_step + 1;

  auto issue_load = [&](int it) {  // producer lane 0 only
    const int tile = first_tile + it * tile_step;
    if (it >= my_tiles) return;
    if (tail_exists && tile == tail_tile) return;  // tail is copied cooperatively
    const int st = it % NSTAGE;
    const uint32_t bar = smem_u32(&full_bar[st]);
    mbar_expect_tx(bar, (uint32_t)(tile_floats * 4 + (VAL_TMA ? TS * A * 4 : 0)));
    bulk_g2s(smem_u32(stage_base) + st * stage_bytes, g_logits + (size_t)tile * tile_floats,
             (uint32_t)(tile_floats * 4), bar);
    if (VAL_TMA)
      bulk_g2s(smem_u32(stage_base) + st * stage_bytes + tile_floats * 4, g_value + (size_t)tile * TS * A,
               (uint32_t)(TS * A * 4), bar);
  };
  // the first tiles are in flight while every thread computes its per-(a,k) constants below
  if (is_producer && lane == 0) {
#pragma unroll
    for (int d = 0; d <= DIST; ++d) issue_load(d);
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
  // element offset (within the row) of this lane's i-th particle, MINUS c: rows are addressed as row_base + c + kofs(i)
  const int spl_c = SPL ? (LPR == 2 ? 15 * c : 12 * (c >> 2)) : 0;
  auto kofs = [&](int i) -> int {
    if (SPL && LPR == 2) return i < 16 ? (spl_c + i) : LPR * i;
    if (SPL && LPR == 8) return 32 * (i >> 2) + 4 * (i & 3) + spl_c;
    return LPR * i;
  };

  constexpr int NCR = CSM ? 1 : EP2;
  float2 isig_r[NCR], nmisig_r[NCR], cst_r[NCR];
  float2 acc1[EP2], acc2[EP2];
  const int tps = A * LPR;               // threads per state slot
  const int tps_idx = tid - slot * tps;  // this thread's column in the constant table
#pragma unroll
  for (int i2 = 0; i2 < EP2; ++i2) {
    float v[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * i2 + h;
      const bool ok = active && (i < EPL) && k_ok(i);
      const int k = c + kofs(i);
      const float ls = ok ? __ldg(&kp.a.logstd[a * P + k]) : 0.f;
      const float mu = ok ? __ldg(&kp.a.loc[a * P + k]) : 0.f;
      const float is = ok ? expf(-ls) : 0.f;
      v[h][0] = is;
      v[h][1] = -mu * is;
      v[h][2] = ok ? -(ls + kHalfLog2Pi) * kLog2e : 0.f;
    }
    if (CSM) {
      if (active && slot == 0) {
        cs[(i2 * 3 + 0) * tps + tps_idx] = make_float2(v[0][0], v[1][0]);
        cs[(i2 * 3 + 1) * tps + tps_idx] = make_float2(v[0][1], v[1][1]);
        cs[(i2 * 3 + 2) * tps + tps_idx] = make_float2(v[0][2], v[1][2]);
      }
    } else {
      isig_r[i2] = make_float2(v[0][0], v[1][0]);
      nmisig_r[i2] = make_float2(v[0][1], v[1][1]);
      cst_r[i2] = make_float2(v[0][2], v[1][2]);
    }
    acc1[i2] = make_float2(0.f, 0.f);
    acc2[i2] = make_float2(0.f, 0.f);
  }
  const float2* csp = cs + (active ? tps_idx : 0);
  auto c_isig = [&](int i2) -> float2 { return CSM ? csp[(i2 * 3 + 0) * tps] : isig_r[CSM ? 0 : i2]; };
  auto c_nmisig = [&](int i2) -> float2 { return CSM ? csp[(i2 * 3 + 1) * tps] : nmisig_r[CSM ? 0 : i2]; };
  auto c_cst = [&](int i2) -> float2 { return CSM ? csp[(i2 * 3 + 2) * tps] : cst_r[CSM ? 0 : i2]; };
  const bool tanh_flag = LEAN ? false : (kp.a.flags & PFPN_HEAD_FLAG_TANH) != 0;
  const bool has_ent_grad = LEAN ? false : kp.has_ent_grad != 0;
  const float eps_clip = kp.a.eps_clip, loss_scale = kp.a.loss_scale;

  float adv_mean = 0.f, adv_rstd = 1.f;
  if (mode == PFPN_HEAD_PPO && kp.a.adv_stats != nullptr) {
    adv_mean = __ldg(&kp.a.adv_stats[0]);
    adv_rstd = __ldg(&kp.a.adv_stats[1]);
  }
  float loss_acc = 0.f;

  __syncthreads();  // mbarrier init visible

  const uint32_t cta_bar_a = smem_u32(cta_bar);
  if (is_producer) {
    // ===================== TMA producer warp ===========================================
    // Mirrors the compute steps: after step it-1 completed (split barrier phase it-1) every
    // compute thread has fenced its gradient STS of tile it-2, so that tile can leave; the
    // stage that load(it+DIST) refills held tile it+DIST-NSTAGE, whose store is older than the
    // NSTAGE-DIST-2 most recent bulk groups.  Only this warp ever blocks on TMA traffic.
    // Iteration `it`: (1) TMA duties that became possible when step it-1 completed, (2) the per-state scalars of tile
    // sit = it-1 (its row partials are complete once step it-1 is).
    for (int it = 1; it <= my_tiles; ++it) {
      const int sit = it - 1;
      const bool do_sip = SIP;
      // the per-state scalars (PPO: adv, lp_old; GRAD: dL/dlp) do not depend on the compute warps: fetch them before
      // blocking on a barrier so the DRAM latency is off the g_bar critical path
      float pf0 = 0.f, pf1 = 0.f;
      if (do_sip && BWD && lane < TS) {
        const int b = (first_tile + sit * tile_step) * TS + lane;
        if (b < B) {
          if (mode == PFPN_HEAD_PPO) {
            pf0 = __ldg(&kp.a.adv[b]);
            pf1 = __ldg(&kp.a.lp_old[b]);
          } else {
            pf0 = __ldg(&kp.a.g_lp[b]);
          }
        }
      }
      if (it >= 1) {
        if (HINT) mbar_wait_hint(cta_bar_a, (uint32_t)((it - 1) & 1), 2000u);
        else mbar_wait(cta_bar_a, (uint32_t)((it - 1) & 1));
        if (lane == 0) {
          if (BWD && it >= 2) {
            const int ptile = first_tile + (it - 2) * tile_step;
            bulk_s2g(g_dlogits + (size_t)ptile * tile_floats,
                     smem_u32(stage_base) + ((it - 2) % NSTAGE) * stage_bytes, (uint32_t)(tile_floats * 4));
            bulk_commit();
            bulk_wait_read<NSTAGE - DIST - 2>();
          }
          issue_load(it + DIST);
        }
        __syncwarp();
      }
      if (do_sip) {
        // ---- per-state scalars of tile sit ----
        const int pb0 = (first_tile + sit * tile_step) * TS;
        if (lane < TS) {
          const int jj = lane / slots, sl = lane - jj * slots;
          const float2* hb = rowbuf + (sit % NSTAGE) * TS * A + jj * (MAXT / GL + 1) + sl * HPS;
          float lp = 0.f, en = 0.f;
#pragma unroll
          for (int h = 0; h < HPS; ++h) {
            const float2 r = hb[h];
            lp += r.x;
            en += r.y;
          }
          const int b = pb0 + lane;
          float g = 0.f;
          if (b < B) {
            kp.a.lp[b] = lp;
            if (kp.a.ent != nullptr) kp.a.ent[b] = en;
            if (BWD) {
              if (mode == PFPN_HEAD_PPO) {
                const float an = (pf0 - adv_mean) * adv_rstd;
                const float ratio = ex2f((lp - pf1) * kLog2e);
                const float surr = ratio * an;
                const float clipped = fminf(fmaxf(ratio, 1.f - eps_clip), 1.f + eps_clip) * an;
                loss_acc -= fminf(surr, clipped) * loss_scale;
                g = (surr <= clipped) ? -loss_scale * ratio * an : 0.f;  // TF Minimum: ties -> x
              } else {
                g = pf0;
              }
            }
          }
          if (BWD) gbuf[(sit % NSTAGE) * TS + lane] = g;
        }
        if (BWD) {
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&g_bar[sit & 1])) : "memory");
        }
      }
    }
  }

  // per-thread fixed row offsets inside a tile and the one "writer" thread per state
  int row_off[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) row_off[j] = (j * slots + slot) * A + a;
  const bool writer = active && a == 0 && c == 0;

  // action / per-state scalar prefetch, one tile ahead.  sc = {adv, lp_old} (PPO) or {g_lp, -}
  float v_nxt[RPT], sc0_nxt[RPT], sc1_nxt[RPT];
  auto load_values = [&](int it) {
    const int tile = first_tile + it * tile_step;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int b = tile * TS + j * slots + slot;
      const bool ok = it < my_tiles && active && b < B;
      v_nxt[j] = (!VAL_TMA && ok) ? __ldg(&g_value[(size_t)b * A + a]) : 0.f;
      if (BWD && !SIP) {
        if (mode == PFPN_HEAD_PPO) {
          sc0_nxt[j] = ok ? __ldg(&kp.a.adv[b]) : 0.f;
          sc1_nxt[j] = ok ? __ldg(&kp.a.lp_old[b]) : 0.f;
        } else {
          sc0_nxt[j] = ok ? __ldg(&kp.a.g_lp[b]) : 0.f;
          sc1_nxt[j] = 0.f;
        }
      } else {
        sc0_nxt[j] = sc1_nxt[j] = 0.f;
      }
    }
  };
  load_values(0);

  const float2 L2 = splat2(kLog2e);
  const float2 nhl = splat2(-0.5f * kLog2e);
  const float2 neg1 = splat2(-1.f);

  // state carried from pass A/B (iteration it) to pass C (iteration it+1); two copies that
  // swap roles every iteration (the loop is unrolled by two so both stay in registers)
  struct RowState {
    float2 e1[RPT][RC ? 1 : EP2], e2[RPT][RC ? 1 : EP2];
    float inv_s1[RPT], s2r[RPT], Hrow[RPT], l2s1[RPT], v_c[RPT], sc0_c[RPT], sc1_c[RPT], nmL_c[RPT];
  };
  RowState stA, stB;

  // One pipeline step: pass A/B of tile `it` into `nw`, then -- after the split-phase CTA
  // barrier of the previous step, arrived on a whole pass A/B ago -- pass C of tile it-1 from `od`.
  auto step = [&](const int it, RowState& nw, RowState& od) {
    if (it > my_tiles) return;
    // ======================= pass A/B of tile it ========================================
    if (it < my_tiles) {
      const int tile = first_tile + it * tile_step;
      const int st = it % NSTAGE;
      float* sbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage_base) + (size_t)st * stage_bytes);
      const bool is_tail = tail_exists && tile == tail_tile;
      const int b0 = tile * TS;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        nw.v_c[j] = v_nxt[j];
        nw.sc0_c[j] = sc0_nxt[j];
        nw.sc1_c[j] = sc1_nxt[j];
      }
      load_values(it + 1);

      if (!is_tail) {
        if (HINT) mbar_wait_hint(smem_u32(&full_bar[st]), (uint32_t)((it / NSTAGE) & 1), 1000u);
        else mbar_wait(smem_u32(&full_bar[st]), (uint32_t)((it / NSTAGE) & 1));
      } else {
        const int nvalid = (B - b0) * AP;
        const float* src = g_logits + (size_t)b0 * AP;
        for (int idx = tid; idx < nvalid; idx += nthr) sbuf[idx] = __ldg(&src[idx]);
        if (VAL_TMA)
          for (int idx = tid; idx < (B - b0) * A; idx += nthr) sbuf[tile_floats + idx] = __ldg(&g_value[(size_t)b0 * A + idx]);
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");  // compute threads only; tail tile only
      }
      if (VAL_TMA) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const bool row_ok = active && (b0 + j * slots + slot < B);
          nw.v_c[j] = sbuf[tile_floats + (row_ok ? row_off[j] : a)];
        }
      }

      float2* rb = rowbuf + (it % NSTAGE) * TS * A;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const bool row_ok = active && (b0 + j * slots + slot < B);
        // masked rows read state 0 of the tile (always valid, finite) and are discarded
        const float* ld = sbuf + (row_ok ? row_off[j] : a) * P + c;
        float2 l[EP2];
        float m = kNegBig;
#pragma unroll
        for (int i2 = 0; i2 < EP2; ++i2) {
          const int i0 = 2 * i2, i1 = 2 * i2 + 1;
          l[i2].x = (i0 < nfull) ? ld[kofs(i0)] : ((i0 == nfull && part_ok) ? ld[kofs(i0)] : kNegBig);
          l[i2].y = (i1 >= EPL) ? kNegBig
                                : ((i1 < nfull) ? ld[kofs(i1)] : ((i1 == nfull && part_ok) ? ld[kofs(i1)] : kNegBig));
          m = max3f(m, l[i2].x, l[i2].y);
        }
        m = row_max<LPR>(m);
        const float2 nmL = splat2(-m * kLog2e);
        const float2 v2 = splat2(nw.v_c[j]);
        float2 s1 = make_float2(0.f, 0.f), s2 = s1, h = s1;
#pragma unroll
        for (int i2 = 0; i2 < EP2; ++i2) {
          const float2 t = fma2(l[i2], L2, nmL);
          float2 x1;
          x1.x = ex2f(t.x);
          x1.y = ex2f(t.y);
          const float2 z = fma2(v2, c_isig(i2), c_nmisig(i2));
          const float2 q = mul2(z, z);
          const float2 u = add2(t, c_cst(i2));
          const float2 t2 = fma2(q, nhl, u);
          float2 x2;
          x2.x = ex2f(t2.x);
          x2.y = ex2f(t2.y);
          s1 = add2(s1, x1);
          s2 = add2(s2, x2);
          h = fma2(x1, t, h);
          if (!RC) {
            nw.e1[j][RC ? 0 : i2] = x1;
            nw.e2[j][RC ? 0 : i2] = x2;
          }
        }
        nw.nmL_c[j] = nmL.x;
        const float S1 = row_sum<LPR>(s1.x + s1.y);
        const float S2 = row_sum<LPR>(s2.x + s2.y);
        const float Hs = row_sum<LPR>(h.x + h.y);
        const float is1 = rcpf(S1);
        const float lg1 = lg2f(S1);
        const float Hval = kLn2 * (lg1 - Hs * is1);
        float lnp = kLn2 * (lg2f(S2) - lg1);  // -inf when every term underflowed (p == 0)
        if (tanh_flag) {
          const float uu = nw.v_c[j];
          lnp -= 2.f * (kLn2 - uu - softplusf_acc(-2.f * uu));
        }
        nw.inv_s1[j] = is1;
        nw.s2r[j] = S2;
        nw.Hrow[j] = Hval;
        nw.l2s1[j] = lg1;
        if (row_ok && c == 0 && g_ent_ba != nullptr) g_ent_ba[(size_t)(b0 + j * slots + slot) * A + a] = Hval;
        if (SEG) {
          // 4 rows of this half-warp -> one partial (lanes differing in bits 2,3 hold different rows)
          float pl = row_ok ? lnp : 0.f, ph = row_ok ? Hval : 0.f;
          pl += __shfl_xor_sync(0xffffffffu, pl, LPR);
          ph += __shfl_xor_sync(0xffffffffu, ph, LPR);
          pl += __shfl_xor_sync(0xffffffffu, pl, 2 * LPR);
          ph += __shfl_xor_sync(0xffffffffu, ph, 2 * LPR);
          if ((tid % GL) == 0) rb[j * (MAXT / GL + 1) + (tid / GL)] = make_float2(pl, ph);
        } else if (row_ok && c == 0) {
          rb[row_off[j]] = make_float2(lnp, Hval);
        }
      }
    }
    // ======================= pass C of tile it-1 (state in registers) ==================
    // Placed first in program order so that the carried registers die before pass A/B
    // allocates the new ones; its barrier phase (it-1) was arrived on an iteration ago.
    if (it >= 1) {
      const int pit = it - 1;
      const int tile = first_tile + pit * tile_step;
      const int st = pit % NSTAGE;
      float* sbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage_base) + (size_t)st * stage_bytes);
      const int b0 = tile * TS;
      const float2* rb = rowbuf + (pit % NSTAGE) * TS * A;

      if (SIP && BWD) {
        if (HINT) mbar_wait_hint(smem_u32(&g_bar[pit & 1]), (uint32_t)((pit >> 1) & 1), 1000u);
        else mbar_wait(smem_u32(&g_bar[pit & 1]), (uint32_t)((pit >> 1) & 1));  // the producer published dL/dlp of tile it-1
      } else {  // (forward: keeps the compute warps within one step of each other and of the producer)
        mbar_wait(cta_bar_a, (uint32_t)(pit & 1));  // every compute thread finished iteration it-1
      }

#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int sidx = j * slots + slot;
        const int b = b0 + sidx;
        const bool row_ok = active && (b < B);
        if (SIP && !BWD) continue;
        // ---- log_prob of the state: sum over a of the per-row log p
        float lp = 0.f, en = 0.f;
        if (SIP) {
          // (done by the producer warp)
        } else if (SEG) {
          const float2* hb = rb + j * (MAXT / GL + 1) + slot * HPS;
#pragma unroll
          for (int h = 0; h < HPS; ++h) {
            const float2 r = hb[h];
            lp += r.x;
            en += r.y;
          }
        } else {
          const float2* rbs = rb + (row_ok ? sidx : 0) * A;
#pragma unroll 3
          for (int aa = c; aa < A; aa += LPR) {
            const float2 r = rbs[aa];
            lp += r.x;
            en += r.y;
          }
          lp = row_sum<LPR>(lp);
          en = row_sum<LPR>(en);
        }
        if (!SIP && writer && row_ok) {
          kp.a.lp[b] = lp;
          if (kp.a.ent != nullptr) kp.a.ent[b] = en;
        }
        if (!BWD) continue;

        float g;
        const float sc0v = od.sc0_c[j];
        if (SIP) {
          g = gbuf[(pit % NSTAGE) * TS + sidx];
        } else if (mode == PFPN_HEAD_PPO) {
          const float sc1v = od.sc1_c[j];
          const float an = (sc0v - adv_mean) * adv_rstd;
          const float ratio = ex2f((lp - sc1v) * kLog2e);
          const float surr = ratio * an;
          const float clipped = fminf(fmaxf(ratio, 1.f - eps_clip), 1.f + eps_clip) * an;
          if (writer && row_ok) loss_acc -= fminf(surr, clipped) * loss_scale;
          g = (surr <= clipped) ? -loss_scale * ratio * an : 0.f;  // TF Minimum: ties -> x
        } else {
          g = sc0v;
        }
        if (!row_ok) g = 0.f;

        float* stp = row_ok ? (sbuf + row_off[j] * P + c) : (dummy + c);
        // RC: same addresses and the same instruction sequence as pass A/B -> bit-identical e1, e2
        const float* ldp = sbuf + (row_ok ? row_off[j] : a) * P + c;
        const float2 nmLc = splat2(od.nmL_c[j]);
        const float2 vRC = splat2(od.v_c[j]);
        auto terms = [&](const int i2, float2& x1, float2& x2) {
          if (!RC) {
            x1 = od.e1[j][RC ? 0 : i2];
            x2 = od.e2[j][RC ? 0 : i2];
            return;
          }
          const int i0 = 2 * i2, i1 = 2 * i2 + 1;
          float2 l;
          l.x = (i0 < nfull) ? ldp[kofs(i0)] : ((i0 == nfull && part_ok) ? ldp[kofs(i0)] : kNegBig);
          l.y = (i1 >= EPL) ? kNegBig
                            : ((i1 < nfull) ? ldp[kofs(i1)] : ((i1 == nfull && part_ok) ? ldp[kofs(i1)] : kNegBig));
          const float2 t = fma2(l, L2, nmLc);
          x1.x = ex2f(t.x);
          x1.y = ex2f(t.y);
          const float2 z = fma2(vRC, c_isig(i2), c_nmisig(i2));
          const float2 q = mul2(z, z);
          const float2 u = add2(t, c_cst(i2));
          const float2 t2 = fma2(q, nhl, u);
          x2.x = ex2f(t2.x);
          x2.y = ex2f(t2.y);
        };
        float dv_rc = 0.f;
        const bool p_ok = od.s2r[j] > 0.f;
        // guard of utils.py:109-117: dL/dp is Inf/NaN when p == 0 -> zeroed
        const float g_row = p_ok ? g : 0.f;
        const float gs2 = p_ok ? g_row * rcpf(od.s2r[j]) : 0.f;
        const float2 v2 = splat2(od.v_c[j]);
        const float2 gs2v = splat2(gs2);
        if (!has_ent_grad) {
          const float2 nc0 = splat2(-g_row * od.inv_s1[j]);
#pragma unroll
          for (int i2 = 0; i2 < EP2; ++i2) {
            float2 e1v, e2v;
            terms(i2, e1v, e2v);
            const float2 rr = mul2(e2v, gs2v);  // g * r_k
            const float2 d = fma2(e1v, nc0, rr);
            const float2 z = fma2(v2, c_isig(i2), c_nmisig(i2));
            const float2 q1 = fma2(z, z, neg1);
            acc1[i2] = fma2(rr, z, acc1[i2]);
            acc2[i2] = fma2(rr, q1, acc2[i2]);
            if (!LEAN && RC && g_dvalue != nullptr) {  // (RC: the logits are gone after this loop)
              const float2 w = mul2(mul2(rr, z), c_isig(i2));
              dv_rc += w.x + w.y;
            }
            const int i0 = 2 * i2, i1 = 2 * i2 + 1;
            if (i0 < nfull) stp[kofs(i0)] = d.x;
            else if (i0 == nfull && part_ok) stp[kofs(i0)] = d.x;
            if (i1 < EPL) {
              if (i1 < nfull) stp[kofs(i1)] = d.y;
              else if (i1 == nfull && part_ok) stp[kofs(i1)] = d.y;
            }
          }
        } else {
          float ge = kp.a.g_ent;
          if (kp.a.g_ent_ba != nullptr && row_ok) ge += __ldg(&kp.a.g_ent_ba[(size_t)b * A + a]);
          // dH/dl_k = -pi_k (ln pi_k + H),  ln pi_k = ln2*(t_k - log2 s1)
          const float2 nc0 = splat2(-(g_row + ge * (od.Hrow[j] - kLn2 * od.l2s1[j])) * od.inv_s1[j]);
          const float2 nc1 = splat2(-ge * kLn2 * od.inv_s1[j]);
#pragma unroll
          for (int i2 = 0; i2 < EP2; ++i2) {
            float2 e1v, e2v;
            terms(i2, e1v, e2v);
            const float2 rr = mul2(e2v, gs2v);
            // t_k = (l_k - m) log2e recovered as log2(e1_k); e1 == 0 contributes nothing
            float2 t;
            t.x = lg2f(fmaxf(e1v.x, 1e-37f));
            t.y = lg2f(fmaxf(e1v.y, 1e-37f));
            const float2 coef = fma2(t, nc1, nc0);
            const float2 d = fma2(e1v, coef, rr);
            const float2 z = fma2(v2, c_isig(i2), c_nmisig(i2));
            const float2 q1 = fma2(z, z, neg1);
            acc1[i2] = fma2(rr, z, acc1[i2]);
            acc2[i2] = fma2(rr, q1, acc2[i2]);
            if (!LEAN && RC && g_dvalue != nullptr) {  // (RC: the logits are gone after this loop)
              const float2 w = mul2(mul2(rr, z), c_isig(i2));
              dv_rc += w.x + w.y;
            }
            const int i0 = 2 * i2, i1 = 2 * i2 + 1;
            if (i0 < nfull) stp[kofs(i0)] = d.x;
            else if (i0 == nfull && part_ok) stp[kofs(i0)] = d.x;
            if (i1 < EPL) {
              if (i1 < nfull) stp[kofs(i1)] = d.y;
              else if (i1 == nfull && part_ok) stp[kofs(i1)] = d.y;
            }
          }
        }
        if (g_dvalue != nullptr) {
          float dv = dv_rc;
          if (!RC) {
#pragma unroll
            for (int i2 = 0; i2 < EP2; ++i2) {
              const float2 rr = mul2(od.e2[j][RC ? 0 : i2], gs2v);
              const float2 z = fma2(v2, c_isig(i2), c_nmisig(i2));
              const float2 w = mul2(mul2(rr, z), c_isig(i2));
              dv += w.x + w.y;
            }
          }
          dv = row_sum<LPR>(dv);
          if (row_ok && c == 0) {
            float out = -dv;
            if (tanh_flag) out += g_row * 2.f * tanhf(od.v_c[j]);
            g_dvalue[(size_t)b * A + a] = out;
          }
        }
      }
      if (BWD) fence_async_smem();  // gradient STS -> visible to the bulk store issued later
    }

    // arrive: this thread's partials of tile it are written, its gradient STS of tile it-1 fenced
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cta_bar_a) : "memory");
  };
  if (!is_producer) {
    for (int it = 0; it <= my_tiles; it += 2) {
      step(it, stA, stB);
      step(it + 1, stB, stA);
    }
  }

  asm volatile("griddepcontrol.launch_dependents;");  // let head_finalize's CTAs get resident behind the drain
  // ---------------- drain: the last tile's gradient ------------------------------------
  mbar_wait(cta_bar_a, (uint32_t)(my_tiles & 1));
  const int prev_tile = my_tiles >= 1 ? first_tile + (my_tiles - 1) * tile_step : -1;
  const int prev_stage = my_tiles >= 1 ? (my_tiles - 1) % NSTAGE : 0;
  const bool prev_was_tail = tail_exists && prev_tile == tail_tile;
  if (BWD && prev_tile >= 0) {
    if (!prev_was_tail) {
      if (is_producer && lane == 0) {
        bulk_s2g(g_dlogits + (size_t)prev_tile * tile_floats, smem_u32(stage_base) + prev_stage * stage_bytes,
                 (uint32_t)(tile_floats * 4));
        bulk_commit();
      }
    } else if (!is_producer) {
      const int b0 = prev_tile * TS;
      const int nvalid = (B - b0) * AP;
      const float* sb = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage_base) + (size_t)prev_stage * stage_bytes);
      float* dst = g_dlogits + (size_t)b0 * AP;
      for (int idx = tid; idx < nvalid; idx += nthr) dst[idx] = sb[idx];
    }
  }
  if (is_producer && lane == 0) bulk_wait_read<0>();  // every bulk group was issued by this lane
  if (BWD) {
    // loss terms live in the writer threads; fixed-order two-level sum
    const float wl = row_sum<32>((is_producer && !SIP) ? 0.f : loss_acc);  // SIP: the producer lanes hold the loss terms
    if (lane == 0) lossbuf[warp] = wl;
    __syncthreads();  // all bulk reads of smem done (the producer waited) before reuse
    float* red = stage_base;  // [slots][2][AP]
    if (active && !is_producer) {
#pragma unroll
      for (int i2 = 0; i2 < EP2; ++i2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 2 * i2 + h;
          if (i < EPL && k_ok(i)) {
            const int k = c + kofs(i);
            red[(slot * 2 + 0) * AP + a * P + k] = h ? acc1[i2].y : acc1[i2].x;
            red[(slot * 2 + 1) * AP + a * P + k] = h ? acc2[i2].y : acc2[i2].x;
          }
        }
      }
    }
    __syncthreads();
    float* part = kp.part + (size_t)blockIdx.x * 2 * AP;
    for (int idx = tid; idx < 2 * AP; idx += nthr + 32) {
      float s = 0.f;
      for (int sl = 0; sl < slots; ++sl) s += red[sl * 2 * AP + idx];
      part[idx] = s;
    }
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w <= nwarps; ++w) s += lossbuf[w];  // (entry nwarps = the producer warp: the SIP loss terms)
      kp.loss_part[blockIdx.x] = s;
    }
  }
}

// Deterministic second stage over the per-CTA partials.
// 32 columns x kFinGroups partial groups per CTA: coalesced 128-byte reads; group g sums parts g, g+G, ... in that
// order and the groups are combined in order, so the result is bit-reproducible.  Every thread issues ALL its loads
// (<= kFinBatch per trip, one trip for <= G*kFinBatch partials) before the first add: one L2 round trip instead of the
// dependent chain the round-1 kernel spent 12 us in.
//
// Optional peer push (N > 1, the sharded head's only exchange): the thread that owns a finished column writes
// it into row `rank` of EVERY rank's gather buffer (remote stores over NVLink), and the last CTA of the grid -- ticket
// counter -- publishes the call number with a system-scope release store into every rank's flag word.  The consumer
// (pfpn_peer_gather_sum) then only reads LOCAL memory.
constexpr int kFinGroups = 32;
constexpr int kFinBatch = 10;
struct HeadPushK {
  float* out[8];   // peer p's gather row for THIS rank ([n] floats), p < nranks
  int* flags[8];   // peer p's flag word for THIS rank
  int* ticket;     // local CTA-arrival counter (self-resetting)
  int nranks;      // 0 = no push
  int value;
  int protocol;             // 1 = 8-byte packets {value, sequence}: see pfpn_head_push
  int consume_value;        // packets: exchange summed by this launch (0 = none)
  const uint2* consume_rows;
  float* consume_out;
  float consume_scale;
};
__device__ __forceinline__ void st_packet(float* row, int i, float v, int seq) {  // one single-copy-atomic 8-byte store
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(row) + i), "r"(__float_as_uint(v)), "r"(seq)
               : "memory");
}
__device__ __forceinline__ uint2 ld_packet(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(32 * kFinGroups) head_finalize_kernel(const float* __restrict__ part,
                                                                        const float* __restrict__ loss_part,
                                                                        const float* __restrict__ logstd, float* __restrict__ dloc,
                                                                        float* __restrict__ dlogstd, float* __restrict__ loss, int AP,
                                                                        int nparts, const HeadPushK push) {
  __shared__ float sh[kFinGroups][33];
  __shared__ int last_cta;
  const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + col;  // column in the [2*AP] partial row
  const bool col_ok = idx < 2 * AP;
  asm volatile("griddepcontrol.launch_dependents;");  // (push mode: lets pfpn_peer_gather_sum get resident and poll early)
  const float ils = (col_ok && idx < AP && grp == 0) ? expf(-logstd[idx]) : 1.f;  // (parameter, not written by the head kernel)
  // packets: the EARLIER exchange this launch sums does not depend on the head kernel -- its loads are issued first, so
  // their latency hides behind the partial sums (and they are not queued behind this launch's remote stores)
  const bool consume = push.protocol == 1 && push.consume_value != 0 && grp == 0 && col_ok;
  const uint2* crow = push.consume_rows + idx;
  const size_t cpitch = 2 * (size_t)AP;
  uint2 pk[8];
#pragma unroll
  for (int r = 0; r < 8; ++r)  // first try: an L2 load (L2 is where peer writes to this GPU's memory land); re-polls are volatile
    pk[r] = (consume && r < push.nranks) ? __ldcg(crow + r * cpitch) : make_uint2(0u, 0u);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the head kernel's partials are complete and visible
  float s = 0.f;
  if (col_ok) {
    for (int p0 = grp; p0 < nparts; p0 += kFinGroups * kFinBatch) {
      float x[kFinBatch];
#pragma unroll
      for (int u = 0; u < kFinBatch; ++u) {
        const int p = p0 + u * kFinGroups;
        x[u] = p < nparts ? __ldcg(&part[(size_t)p * 2 * AP + idx]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kFinBatch; ++u) s += x[u];
    }
  }
  sh[grp][col] = s;
  float l = 0.f;
  const bool loss_warp = loss != nullptr && blockIdx.x == 0 && grp == 1;
  if (loss_warp) {  // one warp sums the per-CTA loss terms (fixed order)
    for (int p0 = col; p0 < nparts; p0 += 32 * kFinBatch) {
      float x[kFinBatch];
#pragma unroll
      for (int u = 0; u < kFinBatch; ++u) {
        const int p = p0 + u * 32;
        x[u] = p < nparts ? __ldcg(&loss_part[p]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kFinBatch; ++u) l += x[u];
    }
  }
  __syncthreads();
  if (grp == 0 && col_ok) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < kFinGroups; ++g) t += sh[g][col];
    t *= ils;
    if (idx < AP) {
      if (dloc != nullptr) dloc[idx] = t;
    } else if (dlogstd != nullptr) {
      dlogstd[idx - AP] = t;
    }
    if (push.protocol == 1) {
      for (int p = 0; p < push.nranks; ++p) st_packet(push.out[p], idx, t, push.value);
      if (consume) {  // the sum of an EARLIER exchange, element idx, rank order
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (r < push.nranks) {
            while ((int)pk[r].y != push.consume_value) pk[r] = ld_packet(crow + r * cpitch);
            sum += __uint_as_float(pk[r].x);
          }
        }
        push.consume_out[idx] = sum * push.consume_scale;
      }
    } else {
      for (int p = 0; p < push.nranks; ++p) push.out[p][idx] = t;
    }
  }
  if (loss_warp) {
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (col == 0) *loss = l;
  }
  if (push.nranks > 0 && push.protocol == 0) {
    // every CTA: my remote stores are ordered before my ticket; the last CTA of the grid raises the flags
    if (grp == 0) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last_cta = (atomicAdd(push.ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last_cta && threadIdx.x < push.nranks) {
      if (threadIdx.x == 0) *push.ticket = 0;  // next call starts from zero again (stream-ordered after this grid)
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(push.flags[threadIdx.x]), "r"(push.value) : "memory");
    }
  }
}

// One CTA: mean and 1/(sqrt(popvar)+1e-8) of adv[B], fixed summation order.
// Eight independent loads per thread per trip keep the single SM's memory pipe busy
// (a dependent one-load loop costs ~20 us at B = 65536, this ~3 us).
__global__ void __launch_bounds__(1024) adv_stats_kernel(const float* __restrict__ adv, int B, float* __restrict__ stats) {
  asm volatile("griddepcontrol.launch_dependents;");  // the head kernel may set up while this one CTA reduces
  __shared__ double sh[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0, ss = 0.0;
  constexpr int U = 8;
  int i = tid;
  for (; i + (U - 1) * 1024 < B; i += U * 1024) {
    float x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = __ldg(&adv[i + u * 1024]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double d = x[u];
      s += d;
      ss += d * d;
    }
  }
  for (; i < B; i += 1024) {
    const double d = __ldg(&adv[i]);
    s += d;
    ss += d * d;
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

// Compiled instantiations.  [pmin, pmax] = particle counts the (LPR, EPL) pair serves; the
// first matching entry whose `kmmask` has the kernel mode's bit is the default for that
// mode, PFPN_HEAD_VARIANT=<n> picks the n-th matching entry instead (tuning aid, read once).
struct HeadVariant {
  int pmin, pmax, lpr, epl, rpt, nstage, maxt, nreg, pt, at, csm;  // csm = OPT bits (1: constants in smem, 2: recompute)
  int kmmask;           // bit KM set: this entry may be the default for kernel mode KM (0 = tuning alternate only)
  head_kernel_t fn[4];  // indexed by KM
};
#define PFPN_HK(LPR, EPL, RPT, NST, KM, MAXT, NREG, PT, AT, OPT) head_kernel<LPR, EPL, RPT, NST, KM, MAXT, NREG, PT, AT, OPT>
#define PFPN_HEAD_VARIANT_ENTRY(PMIN, PMAX, LPR, EPL, RPT, NST, MAXT, NREG, PT, AT, CSM, KMM)                  \
  {                                                                                                            \
    PMIN, PMAX, LPR, EPL, RPT, NST, MAXT, NREG, PT, AT, CSM, KMM, {                                                 \
      PFPN_HK(LPR, EPL, RPT, NST, 0, MAXT, NREG, PT, AT, CSM), PFPN_HK(LPR, EPL, RPT, NST, 1, MAXT, NREG, PT, AT, CSM), \
          PFPN_HK(LPR, EPL, RPT, NST, 2, MAXT, NREG, PT, AT, CSM), PFPN_HK(LPR, EPL, RPT, NST, 3, MAXT, NREG, PT, AT, CSM) \
    }                                                                                                          \
  }
// [pmin, pmax] = particle counts an entry serves.  PT/AT > 0 entries are specialised for
// exactly (P, A) == (PT, AT) -- the shapes BASELINE.json names (A = 36 DeepMimic action
// dims, P in {10, 35, 100}); PT == AT == 0 entries take both at run time.
static const HeadVariant kHeadVariants[] = {
    // P = 35 (DPPO): 2 lanes per row / 18 particles per lane, recompute (OPT bit 1) and bank-conflict-free split
    // ownership (OPT bit 2).  Round-2 measurements at B = 65536 (ms; interleaved ownership in parentheses): PPO .1546
    // (.165), GRAD .1537 (.165; 4 lanes per row with carried terms .179), FWD .0799 (.084), tanh + dvalue .169 (.177).
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 2, 18, 1, 4, 288, 96, 35, 36, 7, 1 | 2 | 4 | 8),
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 2, 18, 1, 4, 288, 96, 35, 36, 3, 0),  // [1] interleaved ownership (round 1 default)
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 2, 18, 1, 4, 288, 96, 35, 36, 15, 0), // [2] split ownership + suspend-time hints (measured: no change, .1544 vs .1545)
    PFPN_HEAD_VARIANT_ENTRY(35, 35, 4, 9, 1, 5, 288, 96, 35, 36, 1, 0),   // [3] 4 lanes per row, carried terms
    // P = 100 (SAC sweep): 13 particles per lane spill in the backward modes when the particle terms are carried,
    // so the backward runs 8 lanes per row WITH recompute (no carried terms, no spills, two CTAs per SM);
    // 16 lanes per row / 7 per lane (one 19-warp CTA per SM) is the carried-terms alternative.  Forward: 8 lanes,
    // constants in registers.  Entries [0] / [1]: the same with split ownership (conflict-free at a 100-float row pitch).
    PFPN_HEAD_VARIANT_ENTRY(100, 100, 8, 13, 1, 4, 288, 96, 100, 36, 7, 2 | 4 | 8),
    PFPN_HEAD_VARIANT_ENTRY(100, 100, 8, 13, 1, 5, 288, 96, 100, 36, 4, 1),
    PFPN_HEAD_VARIANT_ENTRY(100, 100, 8, 13, 1, 4, 288, 96, 100, 36, 3, 0),  // [2] interleaved (round 1 default, backward)
    PFPN_HEAD_VARIANT_ENTRY(100, 100, 8, 13, 1, 5, 288, 96, 100, 36, 0, 0),  // [3] interleaved (round 1 default, forward)
    PFPN_HEAD_VARIANT_ENTRY(100, 100, 16, 7, 1, 6, 576, 96, 100, 36, 1, 0),
    PFPN_HEAD_VARIANT_ENTRY(10, 10, 4, 3, 2, 5, 288, 72, 10, 36, 0, 15),
    PFPN_HEAD_VARIANT_ENTRY(1, 12, 4, 3, 2, 5, 320, 96, 0, 0, 0, 15),
    PFPN_HEAD_VARIANT_ENTRY(13, 36, 4, 9, 1, 5, 288, 96, 0, 0, 0, 15),
    PFPN_HEAD_VARIANT_ENTRY(37, 64, 8, 8, 1, 5, 320, 96, 0, 0, 0, 15),
    PFPN_HEAD_VARIANT_ENTRY(65, 104, 8, 13, 1, 5, 288, 96, 0, 0, 0, 15),
    PFPN_HEAD_VARIANT_ENTRY(105, 256, 16, 16, 1, 4, 320, 168, 0, 0, 0, 15),
};

static const HeadVariant* pick_variant(int A, int P, int km) {
  static const int want = []() {
    const char* e = getenv("PFPN_HEAD_VARIANT");
    return e ? atoi(e) : -1;
  }();
  const HeadVariant* first = nullptr;   // first entry serving (A, P): defines the specialisation group
  const HeadVariant* deflt = nullptr;   // first entry of that group allowed as default for this mode
  int seen = 0;
  for (const HeadVariant& v : kHeadVariants) {
    if (P < v.pmin || P > v.pmax) continue;
    if (v.at != 0 && v.at != A) continue;
    if (first == nullptr) first = &v;
    if (v.pt != first->pt || v.at != first->at) break;
    if (seen == want) return &v;
    if (deflt == nullptr && (v.kmmask >> km & 1)) deflt = &v;
    ++seen;
  }
  return deflt != nullptr ? deflt : first;
}

struct HeadLaunch {
  const HeadVariant* cfg;
  head_kernel_t fn;
  int threads, slots, ts, smem_bytes, ctas_per_sm, num_sms;
};

static int plan_head(int A, int P, int km, HeadLaunch* L) {
  if (A <= 0 || P <= 0) return PFPN_ERR_ARG;
  L->cfg = pick_variant(A, P, km);
  if (L->cfg == nullptr) return PFPN_ERR_UNSUPPORTED;
  const HeadVariant& v = *L->cfg;
  const int per_slot = A * v.lpr;
  if (per_slot > v.maxt) return PFPN_ERR_UNSUPPORTED;
  int slots = v.maxt / per_slot;
  // tiles must start 16-byte aligned: TS*A*P % 4 == 0
  while (slots > 0 && ((slots * v.rpt * A * P) & 3) != 0) --slots;
  if (slots < 1) return PFPN_ERR_UNSUPPORTED;
  if (v.at != 0 && slots != v.maxt / per_slot) return PFPN_ERR_UNSUPPORTED;  // kernel derives slots itself
  L->slots = slots;
  L->ts = slots * v.rpt;
  L->threads = ((slots * per_slot + 31) & ~31) + 32;  // + the TMA producer warp
  const int stage_bytes = (L->ts * A * P * 4 + L->ts * A * 4 + 127) & ~127;  // logits tile + its action values
  L->smem_bytes = v.nstage * stage_bytes + 8 * (v.nstage + 4) + v.nstage * L->ts * A * 8 + (kHeadMaxWarps + 256) * 4 +
                  (v.lpr * v.epl + 1) * 4 + 16 + ((v.csm & 1) ? ((v.epl + 1) / 2) * 3 * per_slot * 8 : 0);
  L->fn = v.fn[km];
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

// Launch with programmatic stream serialization (PDL): the kernel may become resident while its predecessor on the
// stream drains; it orders itself with griddepcontrol.wait before touching global memory.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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
  int rc = plan_head(A, P, mode == PFPN_HEAD_FWD ? 0 : (mode == PFPN_HEAD_PPO ? 2 : 3), &L);
  if (rc != PFPN_OK) return rc;
  out[0] = L.num_sms;
  out[1] = L.ctas_per_sm;
  out[2] = L.threads;
  out[3] = L.ts;
  return PFPN_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int head_logprob_impl(const pfpn_head_args* args, void* workspace, size_t workspace_bytes, const pfpn_head_push* push,
                             pfpn_stream_t stream_) {
  if (args == nullptr) return PFPN_ERR_ARG;
  const pfpn_head_args& a = *args;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a.B < 0 || a.A <= 0 || a.P <= 0) return PFPN_ERR_ARG;
  if (a.mode > PFPN_HEAD_PPO) return PFPN_ERR_ARG;
  if (a.B == 0) return PFPN_OK;
  if (!a.logits || !a.loc || !a.logstd || !a.value || !a.lp) return PFPN_ERR_ARG;
  const bool bwd = a.mode != PFPN_HEAD_FWD;
  if (bwd && !a.dlogits) return PFPN_ERR_ARG;
  if (bwd && push == nullptr && (!a.dloc || !a.dlogstd)) return PFPN_ERR_ARG;  // (with a push the local copies are optional)
  HeadPushK pk;
  memset(&pk, 0, sizeof(pk));
  if (push != nullptr) {
    if (!bwd || push->nranks < 1 || push->nranks > 8 || push->value < 1) return PFPN_ERR_ARG;
    if (push->protocol != 0 && push->protocol != 1) return PFPN_ERR_ARG;
    const bool packets = push->protocol == 1;
    if (!packets && !push->ticket) return PFPN_ERR_ARG;
    for (int p = 0; p < push->nranks; ++p) {
      if (!push->out[p] || (!packets && !push->flags[p])) return PFPN_ERR_ARG;
      if (packets && (reinterpret_cast<uintptr_t>(push->out[p]) & 7)) return PFPN_ERR_ALIGN;
      pk.out[p] = push->out[p];
      pk.flags[p] = push->flags[p];
    }
    pk.ticket = push->ticket;
    pk.nranks = push->nranks;
    pk.value = push->value;
    pk.protocol = push->protocol;
    if (packets && push->consume_value != 0) {
      if (push->consume_value < 1 || push->consume_value >= push->value || !push->consume_rows || !push->consume_out ||
          (reinterpret_cast<uintptr_t>(push->consume_rows) & 7))
        return PFPN_ERR_ARG;
      pk.consume_value = push->consume_value;
      pk.consume_rows = reinterpret_cast<const uint2*>(push->consume_rows);
      pk.consume_out = push->consume_out;
      pk.consume_scale = push->consume_scale;
    }
  }
  if (a.mode == PFPN_HEAD_GRAD && !a.g_lp) return PFPN_ERR_ARG;
  if (a.mode == PFPN_HEAD_PPO && (!a.adv || !a.lp_old || !a.loss)) return PFPN_ERR_ARG;
  if (!aligned16(a.logits) || !aligned16(a.value) || (bwd && !aligned16(a.dlogits))) return PFPN_ERR_ALIGN;

  HeadLaunch L;
  const bool has_ent_grad = (a.g_ent != 0.f || a.g_ent_ba != nullptr);
  const bool lean = !has_ent_grad && !a.dvalue && !a.ent_ba && !(a.flags & PFPN_HEAD_FLAG_TANH);
  const int km = !bwd ? 0 : (!lean ? 1 : (a.mode == PFPN_HEAD_PPO ? 2 : 3));
  int rc = plan_head(a.A, a.P, km, &L);
  if (rc != PFPN_OK) return rc;
  const int num_tiles = (a.B + L.ts - 1) / L.ts;
  int grid = L.num_sms * L.ctas_per_sm;
  if (grid > num_tiles) grid = num_tiles;
  if (grid > kMaxPartCtas) grid = kMaxPartCtas;

  HeadKParams kp;
  kp.a = a;
  kp.num_tiles = num_tiles;
  kp.slots = L.slots;
  kp.has_ent_grad = has_ent_grad ? 1 : 0;
  kp.part = nullptr;
  kp.loss_part = nullptr;
  const size_t AP = (size_t)a.A * a.P;
  if (bwd) {
    const size_t need = (size_t)grid * (2 * AP + 1) * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) return PFPN_ERR_WORKSPACE;
    kp.part = reinterpret_cast<float*>(workspace);
    kp.loss_part = kp.part + (size_t)grid * 2 * AP;
  }
  PFPN_CUDA_OK(launch_pdl(L.fn, dim3(grid), dim3(L.threads), (size_t)L.smem_bytes, stream, kp));
  if (bwd) {
    const int blocks = (int)((2 * AP + 31) / 32);
    PFPN_CUDA_OK(launch_pdl(head_finalize_kernel, dim3(blocks), dim3(32 * kFinGroups), (size_t)0, stream,
                            (const float*)kp.part, (const float*)kp.loss_part, (const float*)a.logstd, a.dloc, a.dlogstd,
                            a.mode == PFPN_HEAD_PPO ? a.loss : (float*)nullptr, (int)AP, grid, pk));
  }
  return PFPN_OK;
}

extern "C" int pfpn_head_logprob(const pfpn_head_args* args, void* workspace, size_t workspace_bytes,
                                 pfpn_stream_t stream_) {
  return head_logprob_impl(args, workspace, workspace_bytes, nullptr, stream_);
}

extern "C" int pfpn_head_logprob_push(const pfpn_head_args* args, void* workspace, size_t workspace_bytes,
                                      const pfpn_head_push* push, pfpn_stream_t stream_) {
  if (push == nullptr) return PFPN_ERR_ARG;
  return head_logprob_impl(args, workspace, workspace_bytes, push, stream_);
}

// Second stage alone, for kernels that produce per-CTA [2*AP] partials in K1's convention (dloc partial = sum g r z,
// divided by sigma here; dlogstd partial as is): csrc/sac_head.cu.
extern "C" int pfpn_head_finalize_partials(const float* part, int32_t nparts, const float* logstd, float* dloc, float* dlogstd,
                                           int32_t AP, pfpn_stream_t stream_) {
  if (!part || !logstd || !dloc || !dlogstd || nparts <= 0 || AP <= 0) return PFPN_ERR_ARG;
  HeadPushK pk;
  memset(&pk, 0, sizeof(pk));
  head_finalize_kernel<<<(2 * AP + 31) / 32, 32 * kFinGroups, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      part, nullptr, logstd, dloc, dlogstd, nullptr, AP, nparts, pk);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}

extern "C" int pfpn_adv_stats(const float* adv, int32_t B, float* stats, pfpn_stream_t stream_) {
  if (!adv || !stats || B <= 0) return PFPN_ERR_ARG;
  adv_stats_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(adv, B, stats);
  PFPN_CUDA_OK(cudaGetLastError());
  return PFPN_OK;
}
