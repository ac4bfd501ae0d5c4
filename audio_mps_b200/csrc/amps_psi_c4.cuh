// Psi scan for bond dimensions 65..128 (DP = 128): one clip per 4-CTA thread-block cluster.
//
// At D = 128 the step matrices N, R, S (128 KB each) no longer fit the register file of one SM, so
// the ROWS of the matrices are split over the CL = 4 CTAs of a cluster: CTA `rank` owns rows
// [32 rank, 32 rank + 32), 16 lanes per row, 8 complex columns of N, R (R^dag), S per thread -- the
// same per-thread shape as the D = 64 kernels.  Every CTA keeps the FULL state vector of the
// current chunk in its own shared memory; each step ends with the owners broadcasting their 32 new
// entries into all four CTAs through distributed shared memory with st.async ... mbarrier::complete_tx
// (data and completion signal in one instruction); consumers wait on their own mbarrier, so neither a
// barrier.cluster nor a __syncthreads sits on the per-step path.  The forward combines the scalars
// that need all rows (E_k = x'^dag S x') once per 16-step chunk by exchanging per-CTA partial sums and
// stores (S x'_k)_i (own rows) and (E_k, |x_k|^2) for the backward, which then needs no exchange.
//
//   model.py:257-267, 276-282, 293-334 (forward), train.py:89 (adjoint) -- same chain form and
//   adjoint as amps_psi.cuh; chunk length CH4 = 16 (rescale / exchange period).
#pragma once
#include "amps_common.cuh"
#include "amps_psi.cuh"

namespace amps {

constexpr int CH4 = 16;

template <int DP, int CL, int NTHREADS = 512>
struct C4 {
  static constexpr int RP = DP / CL;      // rows per CTA
  static constexpr int NTL = NTHREADS;    // threads per CTA (512: 16 lanes per row; 256: 8 lanes, 16 columns each)
  static constexpr int NQ = NTL / RP;     // lanes per row
  static constexpr int CPT = DP / NQ;     // complex columns per thread
  static constexpr int NP = CPT / 2;
  static_assert(CH4 % (NTL / 32) == 0, "whole steps per warp in the chunk-end scalar passes");
  static_assert(RP == 32, "own-row partial sums are one warp wide");
  static_assert(CL <= NQ, "broadcast lanes (the full forward needs 2 CL: checked there)");
};

template <int DP, int CL>
struct alignas(16) FwdC4Smem {
  float2 xs[CH4 + 1][DP];             // x_{k0+kk}, full vector (own rows local, the rest written by peers)
  float2 xps[CH4][DP];                // x'_{k0+kk}, full vector
  float2 qs[2][CH4][DP];              // q_k, double buffered
  float es[CH4][C4<DP, CL>::NTL];     // per-thread partial of Re(x'^dag S x') over this CTA's rows
  float enx[CL][CH4];                 // per-CTA partial sums of every CTA of the cluster
  float2 sps[CH4][C4<DP, CL>::RP];    // (S x'_k)_i for this CTA's rows (kept for the backward)
  float2 evs[CH4];                    // (E_k, |x_k|^2)
  float wav[2][CH4 + 4];
  float sv[2][CH4 + 4];
  float incv[2][CH4];
  double lred[16];
  unsigned long long xbar[2];         // "step s's broadcast has fully arrived", by step parity
};

// grid = CL * (number of clips or virtual clips), cluster = CL, block = 512.
// VIRT: virtual-clip mode of the parallel-in-time scan, as in psi_fwd_uni_kernel.
// SXO: chain only -- x'_k goes where S x'_k would and |x_k|^2 into ev[k].y; S x'_k, E_k and the loss come from
// psi_sx_tc_kernel afterwards (no S mat-vec, no per-chunk exchange of partial sums between the CTAs).
#ifndef AMPS_C4_MINB
#define AMPS_C4_MINB 1
#endif
// OCC2: compiled for two CTAs per SM (64 registers; the chain-only forward needs no more and 103 KB of shared
// memory): with more clips than one-CTA-per-SM clusters fit, two clusters share every SM quadruple and all 148 SMs
// are used (C3: forward chain 134.8 -> 117.7 ms).  Up to that many clips the 99-register build is 5-7 % faster.
template <int DP, int CL, bool VIRT, bool SXO = false, bool OCC2 = false, int NTHREADS = 512>
#if AMPS_C4_MINB
__global__ void __launch_bounds__(NTHREADS, OCC2 ? 2 : 1)
#else
__global__ void __launch_bounds__(NTHREADS)
#endif
    psi_fwd_c4_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                      const float2* __restrict__ matS, const float2* __restrict__ qtab_,
                      const float2* __restrict__ psi0p_, const float* __restrict__ x, int T, AVal A_,
                      float* __restrict__ loss, double* __restrict__ lossd,
                      float2* __restrict__ traj, float* __restrict__ scales, int nchunks,
                      const float2* __restrict__ psi0v, int nvc, int m_steps,
                      float2* __restrict__ sptraj, float2* __restrict__ evout, SegFwd seg) {
  // a kernel queued behind this one with the programmatic-serialisation attribute (launch_waves: the GEMM pass of
  // the clips already done) may start as soon as every CTA of this grid is resident
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const float A = a_get(A_);
  using Cf = C4<DP, CL, NTHREADS>;
  constexpr int NQ = Cf::NQ, CPT = Cf::CPT, RP = Cf::RP, NTL = Cf::NTL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdC4Smem<DP, CL>& sm = *reinterpret_cast<FwdC4Smem<DP, CL>*>(smem_raw);

  const int t = threadIdx.x, il = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int i = (int)rank * RP + il;     // global matrix row of this thread
  int nsteps = T - 1;
  const float* xb = x + (size_t)b * (VIRT ? T : seg.xstride);
  const float2* qtab = qtab_;
  const float2* psi0p = (!VIRT && seg.x0) ? seg.x0 + (size_t)b * seg.x0_stride : psi0p_;
  size_t tstride = T;
  int sstride = nchunks;
  if (VIRT) {
    const int clip = b / nvc, kbeg = (b % nvc) * m_steps;
    nsteps = max(0, min(m_steps, T - 1 - kbeg));
    xb = x + (size_t)clip * T + kbeg;
    qtab = qtab_ + (size_t)kbeg * DP;
    psi0p = psi0v + (size_t)b * DP;
    nchunks = (nsteps + CH4 - 1) / CH4;
    tstride = m_steps + 1;
    sstride = m_steps / CH4;
  }

  float2 Nr[CPT], Rr[CPT], Sr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);
  load_slice<DP, NQ>(Rr, matR, i, jq);
  if (!SXO) load_slice<DP, NQ>(Sr, matS, i, jq);

  if (t < DP) {
    const float2 p = psi0p[t];
    sm.xs[0][t] = p;
    if (traj && rank == 0) traj[(size_t)b * tstride * DP + t] = p;
  }

  auto issue_loads = [&](int c, int buf) {
    const int k0 = c * CH4;
    const int len = min(CH4, nsteps - k0);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[buf][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NTL) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = t; idx <= len; idx += NTL) cp_async4(&sm.wav[buf][idx], xb + k0 + idx);
  };
  auto compute_s = [&](int buf, int len) {
    if (t < len) {
      const float inc = sm.wav[buf][t + 1] - sm.wav[buf][t];   // model.py:263
      sm.incv[buf][t] = inc;
      sm.sv[buf][t] = inc / A;                                  // model.py:303
    }
  };

  // per-step broadcast: lane jq < CL writes x_{k+1,i} into CTA jq, lane CL <= jq < 2CL writes
  // x'_{k,i} into CTA jq - CL (own CTA included: one code path, no local store).
  // Data and completion travel together: st.async ... mbarrier::complete_tx on the TARGET's xbar[s & 1];
  // the consumer waits on its own mbarrier, so no barrier.cluster sits on the per-step critical path.
  // (SXO: x'_k is needed by nobody but the store to global memory -- the owner keeps it locally, only x_{k+1}
  // is broadcast: half the DSMEM traffic per step)
  const bool st_on = jq < (SXO ? CL : 2 * CL), st_x = jq < CL;
  const unsigned st_addr0 = dsmem_addr(st_x ? (const void*)&sm.xs[1][i] : (const void*)&sm.xps[0][i],
                                       (unsigned)(jq & (CL - 1)));
  const unsigned st_bar0 = dsmem_addr(&sm.xbar[0], (unsigned)(jq & (CL - 1)));
  constexpr unsigned STEP_TX = (SXO ? 1 : 2) * DP * sizeof(float2);   // x_{k+1} (and x'_k), all rows
  if (t == 0) {
    mbar_init(&sm.xbar[0], 1);
    mbar_init(&sm.xbar[1], 1);
    mbar_fence_init_cluster();
  }
  int sg = 0;   // steps taken so far by this (virtual) clip: barrier index sg & 1, phase (sg >> 1) & 1

  double lossacc = 0.0;
  if (nchunks > 0) {
    issue_loads(0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    compute_s(0, min(CH4, nsteps));
  }

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    const int k0 = c * CH4;
    const int len = min(CH4, nsteps - k0);
    if (c + 1 < nchunks) issue_loads(c + 1, buf ^ 1);
    cp_async_commit();
    // (D) sv/incv and xs[0] of this chunk visible; EVERY CTA is done with the previous chunk, so its
    // xs / xps / enx may be overwritten remotely from here on
    cluster_sync_all();
    if (!VIRT) if (seg.ckpt && rank == 0 && c % seg.ck_chunks == 0 && t < DP)   // state checkpoint
      seg.ckpt[(size_t)b * seg.ck_stride + (size_t)(c / seg.ck_chunks) * DP + t] = sm.xs[0][t];

    float2 xp_prev = make_float2(0.f, 0.f);
    float s_cur = sm.sv[buf][0];
    // (S x')_i of this CTA's rows, software-pipelined two steps behind the chain: lane 0 of a row ends
    // with the real part, lane 1 with the imaginary part (pair_reduce)
    float2 part_pp = make_float2(0.f, 0.f);
    float* const spf_st = reinterpret_cast<float*>(&sm.sps[0][il]) + (jq & 1);
    const bool ex_on = jq < 2;

    auto step = [&](auto stage_tag, int kk) {
      constexpr int STAGE = decltype(stage_tag)::value;
      constexpr bool FIRST = STAGE == 0;
      constexpr bool EXPC = STAGE == 1 || STAGE == 2;   // expectation pipeline (stage 3: chain only, not first)
      if (t == 0) mbar_arrive_expect_tx(&sm.xbar[sg & 1], STEP_TX);          // arm this step's phase
      if (!FIRST) mbar_wait_cta(&sm.xbar[(sg - 1) & 1], ((sg - 1) >> 1) & 1);    // previous step's rows are in
      float2 xv[CPT], pv[CPT];
#pragma unroll
      for (int m = 0; m < CPT / 2; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.xs[kk][2 * NQ * m + 2 * jq]);
        xv[2 * m] = make_float2(v.x, v.y);
        xv[2 * m + 1] = make_float2(v.z, v.w);
      }
      if (EXPC) {
#pragma unroll
        for (int m = 0; m < CPT / 2; ++m) {
          const float4 v = *reinterpret_cast<const float4*>(&sm.xps[kk - 1][2 * NQ * m + 2 * jq]);
          pv[2 * m] = make_float2(v.x, v.y);
          pv[2 * m + 1] = make_float2(v.z, v.w);
        }
      }
      const float2 q = sm.qs[buf][kk][i];
      const float s_next = sm.sv[buf][kk + 1];
      float red = 0.f;
      if (STAGE == 2) {   // row reduction of step kk-2's partial, level 1
        const bool odd = jq & 1;
        red = (odd ? part_pp.y : part_pp.x) + __shfl_xor_sync(0xffffffffu, odd ? part_pp.x : part_pp.y, 1);
      }
      float2 L[CPT];
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc)
        L[cc] = make_float2(fmaf(s_cur, Rr[cc].x, Nr[cc].x), fmaf(s_cur, Rr[cc].y, Nr[cc].y));
      float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
      for (int cc = 0; cc < CPT; cc += 2) {
        cmac(a0, L[cc], xv[cc]);
        cmac(a1, L[cc + 1], xv[cc + 1]);
      }
      float2 xp = make_float2(a0.x + a1.x, a0.y + a1.y);
      if (STAGE == 2) {
        red += __shfl_xor_sync(0xffffffffu, red, 2);
        red += __shfl_xor_sync(0xffffffffu, red, 4);
      }
      // row reduction over the NQ lanes of a row, the previous step's S x' FMAs in the shuffle shadows
      float2 p0 = make_float2(0.f, 0.f), p1 = p0;
      static_assert(NQ == 16 || SXO, "the expectation pipeline's pair reduction is written for 16 lanes");
      constexpr int LV = NQ == 16 ? 4 : NQ == 8 ? 3 : 2;
      constexpr int CPL = (CPT + LV - 1) / LV;
#pragma unroll
      for (int lv = 0; lv < LV; ++lv) {
        const float ox = __shfl_xor_sync(0xffffffffu, xp.x, 1 << lv);
        const float oy = __shfl_xor_sync(0xffffffffu, xp.y, 1 << lv);
        if (EXPC) {
#pragma unroll
          for (int cc = lv * CPL; cc < (lv + 1) * CPL && cc < CPT; ++cc) {
            if (cc & 1) cmac(p1, Sr[cc], pv[cc]);
            else cmac(p0, Sr[cc], pv[cc]);
          }
        }
        xp.x += ox;
        xp.y += oy;
      }
      const float2 xn = cmul(q, xp);
      st_async_f2_if(st_on, st_addr0 + (unsigned)(kk * DP * (int)sizeof(float2)), st_x ? xn : xp,
                     st_bar0 + (unsigned)((sg & 1) * sizeof(unsigned long long)));
      if (SXO) if (jq == (NQ > CL ? CL : 0)) sm.xps[kk][i] = xp;   // own row of x'_k, local (flushed by this CTA after the chunk)
      if (EXPC) sm.es[kk - 1][t] = fmaf(xp_prev.x, p0.x + p1.x, xp_prev.y * (p0.y + p1.y));
      if (STAGE == 2) {
        red += __shfl_xor_sync(0xffffffffu, red, 8);
        sts_if(ex_on, spf_st + (kk - 2) * (2 * RP), red);
      }
      if (EXPC) part_pp = make_float2(p0.x + p1.x, p0.y + p1.y);
      xp_prev = xp;
      s_cur = s_next;
      ++sg;
    };

    step(IC0{}, 0);
    if (SXO) {
      for (int kk = 1; kk < len; ++kk) step(IC3{}, kk);
    } else {
      if (len > 1) step(IC1{}, 1);
      if (len == CH4) {
#pragma unroll 2
        for (int kk = 2; kk < CH4; ++kk) step(IC2{}, kk);
      } else {
        for (int kk = 2; kk < len; ++kk) step(IC2{}, kk);
      }
    }
    mbar_wait_cta(&sm.xbar[(sg - 1) & 1], ((sg - 1) >> 1) & 1);   // the chunk's last broadcast has landed
    if (!SXO) {  // drain: step len-2's partial is in part_pp; the chunk's last step has none yet
      if (len >= 2) sts_if(ex_on, spf_st + (len - 2) * (2 * RP), pair_reduce<NQ>(part_pp, jq));
      const float2 part = matvec1<DP, NQ>(Sr, sm.xps[len - 1], jq);
      sm.es[len - 1][t] = fmaf(xp_prev.x, part.x, xp_prev.y * part.y);
      sts_if(ex_on, spf_st + (len - 1) * (2 * RP), pair_reduce<NQ>(part, jq));
    }
    cp_async_wait<0>();
    __syncthreads();  // (A)

    if (!SXO) {
    for (int kk = warp; kk < CH4; kk += NTL / 32) {  // this CTA's partial of Re(x'^dag S x') per step, to every CTA of the cluster
        float en = 0.f;
        if (kk < len) {
#pragma unroll
          for (int r = 0; r < NTL / 32; ++r) en += sm.es[kk][lane + 32 * r];
        }
        en = warp_sum_f(en);
        if (lane < CL && kk < len) st_dsmem_f1(dsmem_addr(&sm.enx[rank][kk], (unsigned)lane), en);
      }
      cluster_sync_all();  // (X1)
      for (int kk = warp; kk < CH4; kk += NTL / 32) {  // per-step scalars, identical on every CTA (same operands, same order)
        float nu2 = 0.f;
        if (kk < len) {
#pragma unroll
          for (int r = 0; r < DP / 32; ++r) nu2 += cabs2(sm.xs[kk][lane + 32 * r]);
        }
        nu2 = warp_sum_f(nu2);
        if (lane == 0 && kk < len) {
          float en = 0.f;
#pragma unroll
          for (int r = 0; r < CL; ++r) en += sm.enx[r][kk];
          const float E = en / fmaxf(nu2, 1e-12f);                                  // model.py:324-325 on x'
          const float z = (E * sm.incv[buf][kk]) / A;                // model.py:294
          lossacc -= (double)log1pf(z);
          sm.evs[kk] = make_float2(E, nu2);
        }
      }
    } else {   // |x_k|^2 per step (E_k: psi_sx_tc_kernel)
      for (int kk = warp; kk < CH4; kk += NTL / 32) {
        float nu2 = 0.f;
        if (kk < len) {
#pragma unroll
          for (int r = 0; r < DP / 32; ++r) nu2 += cabs2(sm.xs[kk][lane + 32 * r]);
        }
        nu2 = warp_sum_f(nu2);
        if (lane == 0 && kk < len) sm.evs[kk] = make_float2(0.f, nu2);
      }
    }
    if (c + 1 < nchunks) compute_s(buf ^ 1, min(CH4, nsteps - (k0 + CH4)));
    // rescale by 1/|x_{k0+len}| (every warp of every CTA computes the same norm)
    float n2 = 0.f;
    for (int r = lane; r < DP; r += 32) n2 += cabs2(sm.xs[len][r]);
    n2 = warp_sum_f(n2);
    const float sc = rsqrtf(fmaxf(n2, 1e-12f));   // clamp of model.py:331-333
    __syncthreads();  // every read of xs[0..len] above is done
    if (t < DP) {
      float2 v = sm.xs[len][t];
      v.x *= sc;
      v.y *= sc;
      sm.xs[len][t] = v;
      sm.xs[0][t] = v;
    }
    if (t == 0 && rank == 0 && scales) scales[(size_t)b * sstride + c] = sc;
    if (traj) {
      __syncthreads();  // (B) scaled x_{k0+len} visible
      const float4* src = reinterpret_cast<const float4*>(&sm.xs[1][0]);
      float4* dst = reinterpret_cast<float4*>(traj + ((size_t)b * tstride + k0 + 1) * DP);
      for (int idx = t + (int)rank * NTL; idx < len * DP / 2; idx += NTL * CL) dst[idx] = src[idx];
      if (sptraj) {   // S x'_k (this CTA's rows) and, from rank 0, (E_k, |x_k|^2): row k of the (virtual) clip
        for (int idx = t; idx < len * RP / 2; idx += NTL) {
          const int kk = idx / (RP / 2), r2 = idx % (RP / 2);
          *reinterpret_cast<float4*>(sptraj + ((size_t)b * tstride + k0 + kk) * DP + (int)rank * RP + 2 * r2) =
              SXO ? *reinterpret_cast<const float4*>(&sm.xps[kk][(int)rank * RP + 2 * r2])
                  : *reinterpret_cast<const float4*>(&sm.sps[kk][2 * r2]);
        }
        if (rank == 0 && t < len) evout[(size_t)b * tstride + k0 + t] = sm.evs[t];
      }
    }
  }

  lossacc = warp_sum_d(lossacc);
  if (lane == 0) sm.lred[warp] = lossacc;
  __syncthreads();
  if (t == 0 && rank == 0 && !SXO) {
    double tot = 0.0;
    for (int wv = 0; wv < NTL / 32; ++wv) tot += sm.lred[wv];
    if (loss) loss[b] = (float)tot;
    if (lossd) lossd[b] = tot;
  }
  cluster_sync_all();  // no CTA leaves while a peer could still address its shared memory
}

// CHAIN (the chain-only sweep, TILES = false): no x'_k reconstruction and the inputs double- instead of triple-
// buffered -- 93 KB instead of 163 KB, so that two clusters share every SM quadruple (as in the forward).
template <int DP, int CL, bool CHAIN = false>
struct alignas(16) BwdC4Smem {
  static constexpr int RP = C4<DP, CL>::RP;
  static constexpr int NB = CHAIN ? 2 : 3;
  float2 xs[NB][CH4 + 1][DP];  // trajectory chunk, NB-buffered (3: chunk c, c-1 in use, c-2 landing; 2: c in use, c-1 landing)
  float2 qs[NB][CH4][DP];
  float2 spl[NB][CH4][RP];     // (S x'_k)_i stored by the forward, own rows
  float2 evl[NB][CH4];         // (E_k, |x_k|^2) stored by the forward
  float2 xps[CHAIN ? 1 : 2][CHAIN ? 1 : CH4][DP];   // reconstructed x'_k, full vector (tile fillers only)
  float2 mus[CH4][DP];         // adjoint of x'_k, full vector (own rows local, the rest from peers)
  float wav[NB][CH4 + 4];
  float tt[NB][CH4 + 4];
  float scs[NB][4];
  float sv[2][CH4], incv[2][CH4], dtk[2][CH4], alphas[2][CH4], betas[2][CH4];
  double lred[16];
  unsigned long long mbar[2];  // "mu broadcast number p has fully arrived", by parity of p
};

// Adjoint backward, same recursion and outputs as psi_bwd_uni_kernel; rows split over the cluster.
// TILES = false: chain only -- mu_k goes to row k of mu_out[b] (own rows of every CTA; may alias sptraj, whose
// rows of a chunk are in shared memory two chunks before its mu rows are written) and the gradient tiles are
// contracted afterwards on the tensor cores (amps_tiles_tc.cuh).
template <int DP, int CL, bool VIRT, bool TILES = true, int NTHREADS = 512>
#if AMPS_C4_MINB
__global__ void __launch_bounds__(NTHREADS, 1)
#else
__global__ void __launch_bounds__(NTHREADS)
#endif
    psi_bwd_c4_kernel(const float2* __restrict__ matN, const float2* __restrict__ matRH,
                      const float2* __restrict__ matS, const float2* __restrict__ qtab_,
                      const float* __restrict__ ttab_, const float* __restrict__ x, int T, AVal A_,
                      const float* __restrict__ w, const float2* __restrict__ traj,
                      const float* __restrict__ scales_, int nchunks, float2* __restrict__ Gout,
                      float* __restrict__ gfout, float2* __restrict__ lam0out,
                      double* __restrict__ gAdir, const float2* __restrict__ lam_end, int nvc,
                      int m_steps, const float2* __restrict__ sptraj, const float2* __restrict__ evin,
                      SegBwd seg, float2* __restrict__ mu_out = nullptr) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see psi_fwd_c4_kernel
  const float A = a_get(A_);
  using Cf = C4<DP, CL, NTHREADS>;
  constexpr int NQ = Cf::NQ, CPT = Cf::CPT, NP = Cf::NP, RP = Cf::RP, NTL = Cf::NTL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Sm = BwdC4Smem<DP, CL, !TILES>;
  constexpr int NB = Sm::NB;
  Sm& sm = *reinterpret_cast<Sm*>(smem_raw);

  const int t = threadIdx.x, il = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int i = (int)rank * RP + il;
  int nsteps = T - 1;
  const float* xb = x + (size_t)b * (VIRT ? T : seg.xstride);
  const bool accum = !VIRT && seg.accumulate;
  size_t rows = T;
  const float2* qtab = qtab_;
  const float* ttab = ttab_;
  const float* scales = scales_ + (size_t)b * nchunks;
  float wb;
  if (VIRT) {
    const int clip = b / nvc, kbeg = (b % nvc) * m_steps;
    nsteps = max(0, min(m_steps, T - 1 - kbeg));
    xb = x + (size_t)clip * T + kbeg;
    rows = m_steps + 1;
    qtab = qtab_ + (size_t)kbeg * DP;
    ttab = ttab_ + kbeg;
    scales = scales_ + (size_t)b * (m_steps / CH4);
    nchunks = (nsteps + CH4 - 1) / CH4;
    wb = w[clip];
  } else {
    wb = w[b];
  }
  const float2* trb = traj + (size_t)b * rows * DP;
  const float2* spb = sptraj + (size_t)b * rows * DP + (int)rank * RP;   // this CTA's rows of S x'
  const float2* evb = evin + (size_t)b * rows;
  (void)matS;

  float2 Nr[CPT], Hr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);   // N is Hermitian: N^dag mu uses the same slices
  load_slice<DP, NQ>(Hr, matRH, i, jq);  // R^dag

  float2 GR[CPT], GN[CPT], GE[CPT];
  {
    const float2* Gb = Gout + (size_t)b * 3 * DP * DP;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = Map<DP, NQ>::col(c, jq);
      GR[c] = (TILES && accum) ? Gb[0 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
      GN[c] = (TILES && accum) ? Gb[1 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
      GE[c] = (TILES && accum) ? Gb[2 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
    }
  }

  auto chunk_len = [&](int c) { return min(CH4, nsteps - c * CH4); };

  auto issue_loads = [&](int c) {
    const int lb = c % NB;
    const int k0 = c * CH4;
    const int len = chunk_len(c);
    const float2* xsrc = trb + (size_t)k0 * DP;
    float2* xdst = &sm.xs[lb][0][0];
    for (int idx = t; idx < (len + 1) * DP / 2; idx += NTL) cp_async16(xdst + 2 * idx, xsrc + 2 * idx);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[lb][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NTL) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = t; idx < len * RP / 2; idx += NTL) {
      const int kk = idx / (RP / 2), r2 = idx % (RP / 2);
      cp_async16(&sm.spl[lb][kk][2 * r2], spb + (size_t)(k0 + kk) * DP + 2 * r2);
    }
    for (int idx = t; idx <= len; idx += NTL) {
      cp_async4(&sm.wav[lb][idx], xb + k0 + idx);
      cp_async4(&sm.tt[lb][idx], ttab + k0 + idx);
    }
    for (int idx = t; idx < len; idx += NTL) cp_async8(&sm.evl[lb][idx], evb + k0 + idx);
    if (t == 0) cp_async4(&sm.scs[lb][0], scales + c);
  };

  double gAacc = 0.0;
  // s, inc, dt, alpha_k, beta_k, the direct dL/dA term (from the forward's (E_k, |x_k|^2));
  // x'_k = conj(q_k) x_{k+1} / c_k (full vector, every CTA)
  auto prep_elementwise = [&](int c) {
    const int lb = c % NB, ds = c & 1, len = chunk_len(c);
    if (t < len) {
      const float inc = sm.wav[lb][t + 1] - sm.wav[lb][t];
      const float s = inc / A;
      sm.incv[ds][t] = inc;
      sm.sv[ds][t] = s;
      sm.dtk[ds][t] = sm.tt[lb][t + 1] - sm.tt[lb][t];
      const float2 ev = sm.evl[lb][t];
      const float E = ev.x, nu2 = ev.y;
      const float arg = 1.0f + (E * inc) / A;
      const float gE = wb * (-s / arg);
      const float alpha = 2.0f * gE / fmaxf(nu2, 1e-12f);
      sm.alphas[ds][t] = alpha;
      sm.betas[ds][t] = -alpha * E;
      gAacc += (double)wb * (double)E * (double)inc / ((double)A * (double)A * (double)arg);
      // chain-only sweep: (alpha_k, 1/c_k) replace the consumed (E_k, |x_k|^2) for the tile kernel
      if (!TILES) if (mu_out && rank == 0)
        const_cast<float2*>(evb)[(size_t)c * CH4 + t] = make_float2(alpha, t == len - 1 ? 1.0f / sm.scs[lb][0] : 1.0f);
    }
    if (TILES) {   // x'_k only feeds the G_E tile
      const float inv_sc = 1.0f / sm.scs[lb][0];
      for (int idx = t; idx < len * DP; idx += NTL) {
        const int kk = idx / DP, r = idx % DP;
        float2 xp = cmul_ca(sm.qs[lb][kk][r], sm.xs[lb][kk + 1][r]);
        if (kk == len - 1) {
          xp.x *= inv_sc;
          xp.y *= inv_sc;
        }
        sm.xps[ds][kk][r] = xp;
      }
    }
  };
  float2 lam = make_float2(0.f, 0.f);  // adjoint of x_{k+1,i}, replicated over the NQ lanes
  if (VIRT) {
    if (lam_end) lam = lam_end[(size_t)b * DP + i];
  } else if (seg.lam_end) {
    lam = seg.lam_end[(size_t)b * DP + i];
  }
  float gf = accum ? gfout[(size_t)b * DP + i] : 0.f;
  if (t == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    mbar_fence_init_cluster();
  }

  if (nchunks > 0) {
    const int cl = nchunks - 1;
    issue_loads(cl);
    cp_async_commit();
    if (NB == 3) {
      if (cl >= 1) issue_loads(cl - 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      prep_elementwise(cl);
    }
    cluster_sync_all();   // every CTA of the cluster has started: remote shared memory is addressable
  }

  // lanes jq < CL broadcast mu_{k,i} into CTA jq: st.async completing on the target's mbar[p & 1]
  const bool mu_on = jq < CL;
  const unsigned mu_addr0 = dsmem_addr(&sm.mus[0][i], (unsigned)(jq & (CL - 1)));
  const unsigned mu_bar0 = dsmem_addr(&sm.mbar[0], (unsigned)(jq & (CL - 1)));
  constexpr unsigned MU_ROW = DP * sizeof(float2);
  int pg = 0;   // mu broadcasts so far: barrier index pg & 1, phase (pg >> 1) & 1

  for (int c = nchunks - 1; c >= 0; --c) {
    const int lb = c % NB, ds = c & 1;
    const int len = chunk_len(c);
    if (NB == 3) {
      if (c >= 2) issue_loads(c - 2);
      cp_async_commit();
      cp_async_wait<1>();   // chunk c-1 has landed
      __syncthreads();      // (T1) alphas/betas of chunk c visible
      if (c >= 1) prep_elementwise(c - 1);
    } else {
      cp_async_wait<0>();   // chunk c has landed (issued one chunk = ~15 k cycles ago)
      __syncthreads();      // ... for every thread, and everybody is done with chunk c+1's buffer
      prep_elementwise(c);
      if (c >= 1) issue_loads(c - 1);   // into the buffer chunk c+1 has just left
      cp_async_commit();
      __syncthreads();      // (T1) alphas/betas of chunk c visible
    }
    const float sc = sm.scs[lb][0];

    float2 mu;
    {  // adjoint of x' for the chunk's last step (carries the rescale c_k)
      const int kk = len - 1;
      const float2 xn = sm.xs[lb][kk + 1][i];
      gf = fmaf(sm.dtk[ds][kk], lam.x * xn.y - lam.y * xn.x, gf);   // Im(conj(lam) x_{k+1})
      mu = cmul_ca(sm.qs[lb][kk][i], lam);
      const float al = sm.alphas[ds][kk];
      const float2 sp = sm.spl[lb][kk][il];
      mu.x = fmaf(al, sp.x, mu.x * sc);
      mu.y = fmaf(al, sp.y, mu.y * sc);
      if (t == 0) mbar_arrive_expect_tx(&sm.mbar[pg & 1], MU_ROW);
      st_async_f2_if(mu_on, mu_addr0 + (unsigned)kk * MU_ROW, mu,
                     mu_bar0 + (unsigned)((pg & 1) * sizeof(unsigned long long)));
      ++pg;
    }
    __syncthreads();      // (T2) x' of chunk c-1 visible (the mu broadcast is awaited per step)

    auto step = [&](int kk) {
      if (t == 0 && kk > 0) mbar_arrive_expect_tx(&sm.mbar[pg & 1], MU_ROW);   // arm this step's broadcast
      mbar_wait_cta(&sm.mbar[(pg - 1) & 1], ((pg - 1) >> 1) & 1);                  // mu_kk is in
      float2 mv[CPT];
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.mus[kk][2 * NQ * m + 2 * jq]);
        mv[2 * m] = make_float2(v.x, v.y);
        mv[2 * m + 1] = make_float2(v.z, v.w);
      }
      const float s = sm.sv[ds][kk];
      const float be = sm.betas[ds][kk];
      const float2 xk = sm.xs[lb][kk][i];
      const int km = kk > 0 ? kk - 1 : 0;
      const float2 q1 = sm.qs[lb][km][i];
      const float al1 = sm.alphas[ds][km];
      const float2 sp1 = sm.spl[lb][km][il];
      const float dt1 = sm.dtk[ds][km];
      // ---- chain: lam_i = (L_k^dag mu)_i + beta x_{k,i} ----------------------------------
      float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
      for (int cc = 0; cc < CPT; cc += 2) {
        const float2 l0 = make_float2(fmaf(s, Hr[cc].x, Nr[cc].x), fmaf(s, Hr[cc].y, Nr[cc].y));
        const float2 l1 = make_float2(fmaf(s, Hr[cc + 1].x, Nr[cc + 1].x), fmaf(s, Hr[cc + 1].y, Nr[cc + 1].y));
        cmac(a0, l0, mv[cc]);
        cmac(a1, l1, mv[cc + 1]);
      }
      float2 lp = make_float2(a0.x + a1.x, a0.y + a1.y);
      float ox = __shfl_xor_sync(0xffffffffu, lp.x, 1);
      float oy = __shfl_xor_sync(0xffffffffu, lp.y, 1);
      // ---- filler 1: rank-1 tiles of step kk (rows of this CTA) -------------------------
      if (TILES) {
        const float2 xpi = sm.xps[ds][kk][i];
        const float al = sm.alphas[ds][kk];
        const float2 u1 = make_float2(s * mu.x, s * mu.y);
        const float2 u3 = make_float2(al * xpi.x, al * xpi.y);
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(&sm.xs[lb][kk][2 * NQ * m + 2 * jq]);
          const float4 pv = *reinterpret_cast<const float4*>(&sm.xps[ds][kk][2 * NQ * m + 2 * jq]);
          const float2 x0 = make_float2(xv.x, xv.y), x1 = make_float2(xv.z, xv.w);
          const float2 p0 = make_float2(pv.x, pv.y), p1 = make_float2(pv.z, pv.w);
          cmac_cx(GR[2 * m], u1, x0);
          cmac_cx(GR[2 * m + 1], u1, x1);
          cmac_cx(GN[2 * m], mu, x0);
          cmac_cx(GN[2 * m + 1], mu, x1);
          cmac_cx(GE[2 * m], u3, p0);
          cmac_cx(GE[2 * m + 1], u3, p1);
        }
      }
      lp.x += ox;
      lp.y += oy;
      ox = __shfl_xor_sync(0xffffffffu, lp.x, 2);
      oy = __shfl_xor_sync(0xffffffffu, lp.y, 2);
      lp.x += ox;
      lp.y += oy;
      if (NQ >= 8) {
        lp.x += __shfl_xor_sync(0xffffffffu, lp.x, 4);
        lp.y += __shfl_xor_sync(0xffffffffu, lp.y, 4);
      }
      if (NQ == 16) {
        lp.x += __shfl_xor_sync(0xffffffffu, lp.x, 8);
        lp.y += __shfl_xor_sync(0xffffffffu, lp.y, 8);
      }
      static_assert(NQ == 16 || NQ == 8 || NQ == 4, "row reduction levels");
      lam.x = fmaf(be, xk.x, lp.x);
      lam.y = fmaf(be, xk.y, lp.y);
      // ---- adjoint of x' for step kk-1 --------------------------------------------------
      if (kk > 0) {
        gf = fmaf(dt1, lam.x * xk.y - lam.y * xk.x, gf);   // Im(conj(lam) x_k), x_k = x_{(k-1)+1}
        mu = cmul_ca(q1, lam);
        mu.x = fmaf(al1, sp1.x, mu.x);
        mu.y = fmaf(al1, sp1.y, mu.y);
      }
      st_async_f2_if(mu_on && kk > 0, mu_addr0 + (unsigned)km * MU_ROW, mu,
                     mu_bar0 + (unsigned)((pg & 1) * sizeof(unsigned long long)));
      if (kk > 0) ++pg;
    };

    for (int kk = len - 1; kk >= 0; --kk) step(kk);
    if (!TILES) if (mu_out) {
      // flush this CTA's rows of the chunk's mu ring (all of them have arrived: every step waited for its
      // broadcast; a peer that is already one broadcast ahead only touches ITS rows of the ring)
      float2* dst = mu_out + ((size_t)b * rows + (size_t)c * CH4) * DP + (int)rank * RP;
      for (int idx = t; idx < len * RP; idx += NTL) {
        const int kk = idx / RP, r = idx % RP;
        dst[(size_t)kk * DP + r] = sm.mus[kk][(int)rank * RP + r];
      }
    }
  }
  cp_async_wait<0>();

  // ---- per-clip outputs (rows of this CTA) ---------------------------------------------------
  if (TILES) {
    float2* Gb = Gout + (size_t)b * 3 * DP * DP;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = Map<DP, NQ>::col(c, jq);
      Gb[0 * DP * DP + i * DP + col] = GR[c];
      Gb[1 * DP * DP + i * DP + col] = GN[c];
      Gb[2 * DP * DP + i * DP + col] = GE[c];
    }
  }
  if (jq == 0) {
    gfout[(size_t)b * DP + i] = gf;
    lam0out[(size_t)b * DP + i] = lam;
  }
  gAacc = warp_sum_d(gAacc);
  if (lane == 0) sm.lred[warp] = gAacc;
  __syncthreads();
  if (t == 0 && rank == 0) {
    double tot = accum ? gAdir[b] : 0.0;
    for (int wv = 0; wv < NTL / 32; ++wv) tot += sm.lred[wv];
    gAdir[b] = tot;
  }
  cluster_sync_all();
}

// -------------------------------------------------------------------------------------------
// Sampler for bond dimensions 65..128 (model.py:242-251, 284-291): one waveform per 4-CTA cluster,
// rows of N and R split as above.  The sampler feeds E(psi_k) back into the increment, so a step needs
// every row's <x, R x> and |x|^2 BEFORE the new state exists -- two all-to-all exchanges per step if the
// owners formed x_{k+1} themselves.  Here the owners broadcast the two mat-vec results a_i = (N x_k)_i,
// y_i = (R x_k)_i (one 16-byte st.async per row and target CTA) and every CTA its partial
// (sum conj(x_i) y_i, sum |x_i|^2) into the SAME mbarrier phase; after that single wait every thread has
// s_k and the norm and forms the eight entries  x_{k+1,j} = c q_{k,j} (a_j + s_k y_j)  its next mat-vec
// reads straight into registers (the state itself is never stored).  One intra-CTA barrier (partial
// sums) and one cluster exchange per step.
// -------------------------------------------------------------------------------------------
template <int DP, int CL>
struct alignas(16) SampleC4Smem {
  float4 ay[2][DP];              // (a_j, y_j) of every row, by step parity
  float2 part[2][CL];            // per-CTA (sum conj(x_i) y_i, sum |x_i|^2), by step parity
  float2 qs[2][CH4][DP];         // q_k, double buffered by chunk
  float nz[2][CH4];
  float outs[CH4];
  float2 wred[2][16];
  unsigned long long xbar[2];
};

// grid = CL * n, cluster = CL, block = 512
template <int DP, int CL>
__global__ void __launch_bounds__(512)
    psi_sample_c4_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                         const float2* __restrict__ qtab, const float2* __restrict__ psi0p,
                         const float* __restrict__ noise, int L, int n, float A, float dtf,
                         float* __restrict__ out) {
  using Cf = C4<DP, CL>;
  constexpr int NQ = Cf::NQ, CPT = Cf::CPT, RP = Cf::RP, NTL = Cf::NTL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SampleC4Smem<DP, CL>& sm = *reinterpret_cast<SampleC4Smem<DP, CL>*>(smem_raw);
  const int t = threadIdx.x, il = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int i = (int)rank * RP + il;
  const int nchunks = (L + CH4 - 1) / CH4;

  float2 Nr[CPT], Rr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);
  load_slice<DP, NQ>(Rr, matR, i, jq);

  auto issue_loads = [&](int c, int buf) {
    const int k0 = c * CH4, len = min(CH4, L - k0);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[buf][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NTL) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = t; idx < len; idx += NTL) cp_async4(&sm.nz[buf][idx], noise + (size_t)b * L + k0 + idx);   // [n][L]
  };
  if (t == 0) {
    mbar_init(&sm.xbar[0], 1);
    mbar_init(&sm.xbar[1], 1);
    mbar_fence_init_cluster();
  }
  if (nchunks > 0) issue_loads(0, 0);
  cp_async_commit();
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers are initialised before any remote st.async

  constexpr unsigned STEP_TX = DP * sizeof(float4) + CL * sizeof(float2);
  const bool bc_on = jq < CL;                              // lane jq < CL sends (a_i, y_i) to CTA jq
  const unsigned ay_addr0 = dsmem_addr(&sm.ay[0][i], (unsigned)(jq & (CL - 1)));
  const unsigned bar_addr0 = dsmem_addr(&sm.xbar[0], (unsigned)(jq & (CL - 1)));
  const unsigned pt_addr0 = dsmem_addr(&sm.part[0][rank], (unsigned)(t & (CL - 1)));   // (threads t < CL)
  const unsigned pb_addr0 = dsmem_addr(&sm.xbar[0], (unsigned)(t & (CL - 1)));

  // mat-vec of the state held in registers (this thread's 8 columns) + exchange of step `sg`
  float2 a_own = make_float2(0.f, 0.f), y_own = a_own, x_own = a_own;
  int sg = 0;
  auto matvec_and_send = [&](const float2 (&xv)[CPT]) {
    const int p = sg & 1;
    if (t == 0) mbar_arrive_expect_tx(&sm.xbar[p], STEP_TX);
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
    for (int cc = 0; cc < CPT; cc += 2) {
      cmac(a0, Nr[cc], xv[cc]);
      cmac(a1, Nr[cc + 1], xv[cc + 1]);
      cmac(b0, Rr[cc], xv[cc]);
      cmac(b1, Rr[cc + 1], xv[cc + 1]);
    }
    float4 ayv = make_float4(a0.x + a1.x, a0.y + a1.y, b0.x + b1.x, b0.y + b1.y);
#pragma unroll
    for (int m = 1; m < NQ; m <<= 1) {
      ayv.x += __shfl_xor_sync(0xffffffffu, ayv.x, m);
      ayv.y += __shfl_xor_sync(0xffffffffu, ayv.y, m);
      ayv.z += __shfl_xor_sync(0xffffffffu, ayv.z, m);
      ayv.w += __shfl_xor_sync(0xffffffffu, ayv.w, m);
    }
    a_own = make_float2(ayv.x, ayv.y);
    y_own = make_float2(ayv.z, ayv.w);
    if (bc_on) {
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(
                       ay_addr0 + (unsigned)(p * DP * sizeof(float4))),
                   "f"(ayv.x), "f"(ayv.y), "f"(ayv.z), "f"(ayv.w), "r"(bar_addr0 + (unsigned)(p * sizeof(unsigned long long)))
                   : "memory");
    }
    // this CTA's partial of <x, R x> and |x|^2 (values replicated over a row's 16 lanes: two rows per warp)
    float e = fmaf(x_own.x, y_own.x, x_own.y * y_own.y), nn = cabs2(x_own);
    e += __shfl_xor_sync(0xffffffffu, e, 16);
    nn += __shfl_xor_sync(0xffffffffu, nn, 16);
    if (lane == 0) sm.wred[p][warp] = make_float2(e, nn);
    __syncthreads();
    if (t < CL) {
      float es = 0.f, ns = 0.f;
#pragma unroll
      for (int wv = 0; wv < NTL / 32; ++wv) {
        es += sm.wred[p][wv].x;
        ns += sm.wred[p][wv].y;
      }
      st_async_f2_if(true, pt_addr0 + (unsigned)(p * CL * sizeof(float2)), make_float2(es, ns),
                     pb_addr0 + (unsigned)(p * sizeof(unsigned long long)));
    }
    ++sg;
  };

  // step 0's mat-vec: x_0 = psi_0
  {
    float2 xv[CPT];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) xv[cc] = psi0p[Map<DP, NQ>::col(cc, jq)];
    x_own = psi0p[i];
    if (L > 0) matvec_and_send(xv);
  }
  float X = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1, k0 = c * CH4, len = min(CH4, L - k0);
    if (c + 1 < nchunks) issue_loads(c + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();   // q_k / noise of this chunk visible; outs flushed
    for (int kk = 0; kk < len; ++kk) {
      const int p = (sg - 1) & 1;
      mbar_wait_cta(&sm.xbar[p], ((sg - 1) >> 1) & 1);     // (a, y) of every row and the four partials are in
      float es = 0.f, nsum = 0.f;
#pragma unroll
      for (int r = 0; r < CL; ++r) {
        const float2 pr = sm.part[p][r];
        es += pr.x;
        nsum += pr.y;
      }
      const float E = 2.0f * es / fmaxf(nsum, 1e-12f);                         // model.py:319-325
      const float inc = __fadd_rn(__fmul_rn(E, dtf), sm.nz[buf][kk]);           // model.py:286
      X = __fadd_rn(X, inc);                                                    // model.py:287
      const float s = inc / A;                                                  // model.py:303
      const float rn = rsqrtf(fmaxf(nsum, 1e-12f));   // lagged normalisation keeps |x| ~ 1
      if (t == 0 && rank == 0) sm.outs[kk] = A * X;                             // model.py:251
      const bool more = k0 + kk + 1 < L;
      // x_{k+1}: own row from registers, the eight mat-vec columns from the broadcast vectors
      {
        const float2 xp = make_float2(fmaf(s, y_own.x, a_own.x) * rn, fmaf(s, y_own.y, a_own.y) * rn);
        x_own = cmul(sm.qs[buf][kk][i], xp);
      }
      if (more) {
        float2 xv[CPT];
#pragma unroll
        for (int cc = 0; cc < CPT; ++cc) {
          const int col = Map<DP, NQ>::col(cc, jq);
          const float4 v = sm.ay[p][col];
          const float2 xp = make_float2(fmaf(s, v.z, v.x) * rn, fmaf(s, v.w, v.y) * rn);
          xv[cc] = cmul(sm.qs[buf][kk][col], xp);
        }
        matvec_and_send(xv);
      }
    }
    __syncthreads();
    if (rank == 0 && t < len) out[(size_t)b * L + k0 + t] = sm.outs[t];
  }
  cp_async_wait<0>();
  cluster_sync_all();   // no CTA leaves while a peer could still address its shared memory
}

}  // namespace amps
