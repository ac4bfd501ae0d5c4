// Cluster variants of the Psi scan kernels (sm_100a thread-block clusters + distributed shared
// memory).  Used when the batch leaves SMs idle (2*B <= #SMs, e.g. BASELINE config C1: 64 clips on
// 148 SMs): each clip gets a CLUSTER OF TWO CTAs on two SMs --
//     rank 0  CHAIN CTA : nothing but the sequential recursion (4 chain warps + 1 loader warp)
//     rank 1  FILLER CTA: everything else, on its own SM, so it no longer competes with the chain
//                         for FMA issue slots or the shared-memory pipe
// The filler pulls each finished 32-step chunk (x_k, x'_k rings) out of the chain CTA's shared
// memory with ld.shared::cluster, and (backward) pushes the packed per-row chain inputs into it with
// st.shared::cluster; the two CTAs meet at one barrier.cluster per chunk.
// Measured motivation (profiles/): with the filler warps on the same SM the forward chain runs at
// ~410 cycles/step, alone at ~297.
#pragma once
#include "amps_psi.cuh"

namespace amps {

template <int DP, int NQ>
struct alignas(16) FwdClSmem {
  static constexpr int NTC = DP * NQ, G = NTC / CH, ES = NTC + (G < 32 ? G : 0);
  float2 xs[2][CH + 1][DP];   // chain CTA: rings by chunk parity; filler CTA: xs[0] = local copy
  float2 xps[2][CH][DP];
  float2 qs[2][CH][DP];       // chain CTA
  float es[CH][ES];           // filler CTA
  float2 spp[CH][NQ][DP];     // filler CTA: per-lane partials of S x'_k
  float wav[2][CH + 4];
  float sv[2][CH + 4];
  float incv[2][CH];
  double lred[32];
  float scv[2][4];                   // chain CTA: rescale factor of the chunk in ring p
  unsigned long long full_bar[2];    // filler CTA: ring p of the chain CTA holds a finished chunk
  unsigned long long empty_bar[2];   // chain CTA : the filler has pulled ring p
};

// block = 2*NTC + 32 threads; grid = 2*B CTAs in clusters of 2.  The block size is set by the FILLER CTA: with
// one work group of NTC threads its per-chunk work (pull, S x', E_k, loss, trajectory / S x' / (E, nu^2) stores)
// took longer than the chain's 32 steps when everything is kept for the backward (308 instead of 280
// cycles per step); two work groups share the steps of a chunk.  In the chain CTA the extra warps only
// attend the per-chunk __syncthreads.
template <int DP, int NQ>
__global__ void __launch_bounds__(2 * DP * NQ + 32, 1)
    psi_fwd_cl_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                      const float2* __restrict__ matS, const float2* __restrict__ qtab,
                      const float2* __restrict__ psi0p, const float* __restrict__ x, int T, AVal A_,
                      float* __restrict__ loss, double* __restrict__ lossd,
                      float2* __restrict__ traj, float* __restrict__ scales, int nchunks,
                      float2* __restrict__ sptraj, float2* __restrict__ evout, SegFwd seg) {
  const float A = a_get(A_);
  using M = Map<DP, NQ>;
  using Sm = FwdClSmem<DP, NQ>;
  constexpr int NTC = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int G = Sm::G;
  constexpr int LV = (NQ == 4) ? 2 : 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm& sm = *reinterpret_cast<Sm*>(smem_raw);

  const int t = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x >> 1;
  const int nsteps = T - 1;
  const float* xb = x + (size_t)b * seg.xstride;
  const int lane = t & 31;
  auto chunk_len = [&](int c) { return min(CH, nsteps - c * CH); };

  if (t == 0) {
    mbar_init(&sm.full_bar[0], 1);
    mbar_init(&sm.full_bar[1], 1);
    mbar_init(&sm.empty_bar[0], 1);
    mbar_init(&sm.empty_bar[1], 1);
    mbar_fence_init_cluster();
  }
  __syncthreads();
  cluster_arrive_release();   // both CTAs' barriers are initialised before any remote arrive
  cluster_wait_acquire();

  if (rank == 0) {
    // ===================================== CHAIN CTA ==========================================
    const bool is_loader = t >= NTC && t < NTC + 32;   // one warp: prefetch of q_k / waveform, s_k
    const bool is_chain = t < NTC;
    const int i = t / NQ, jq = t % NQ;         // (chain threads)
    auto load_inputs = [&](int c) {            // loader warp only
      const int buf = c & 1, k0 = c * CH, len = chunk_len(c);
      const float2* qsrc = qtab + (size_t)k0 * DP;
      float2* qdst = &sm.qs[buf][0][0];
      for (int idx = lane; idx < len * DP / 2; idx += 32) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
      for (int idx = lane; idx <= len; idx += 32) cp_async4(&sm.wav[buf][idx], xb + k0 + idx);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      if (lane < len) sm.sv[buf][lane] = (sm.wav[buf][lane + 1] - sm.wav[buf][lane]) / A;   // model.py:263,303
    };
    if (t < DP) {
      const float2 p = seg.x0 ? seg.x0[(size_t)b * seg.x0_stride + t] : psi0p[t];
      sm.xs[0][0][t] = p;
      if (traj) traj[(size_t)b * T * DP + t] = p;
    }
    if (is_loader && nchunks > 0) load_inputs(0);
    __syncthreads();

    float2 Nr[CPT], Rr[CPT];
    if (is_chain) {
      load_slice<DP, NQ>(Nr, matN, i, jq);
      load_slice<DP, NQ>(Rr, matR, i, jq);
    }
    float2 vstart = make_float2(0.f, 0.f);
    const unsigned rfull = dsmem_addr(&sm.full_bar[0], 1);
    for (int c = 0; c < nchunks; ++c) {
      {
        const int p = c & 1, len = chunk_len(c);
        if (is_loader) {
          if (c + 1 < nchunks) load_inputs(c + 1);
        } else if (is_chain) {
          float2* const st2 = (jq == 0) ? &sm.xs[p][1][i] : &sm.xps[p][0][i];
          const bool st2_on = jq < 2;
          float s_cur = sm.sv[p][0];
          const unsigned xs_addr = smem_addr_pinned(&sm.xs[p][0][2 * jq]);
          auto step = [&](int kk) {
            float2 xv[CPT];
#pragma unroll
            for (int m = 0; m < CPT / 2; ++m) {
              const float4 v = lds128v(xs_addr + (unsigned)((kk * DP + 2 * NQ * m) * sizeof(float2)));
              xv[2 * m] = make_float2(v.x, v.y);
              xv[2 * m + 1] = make_float2(v.z, v.w);
            }
            const float2 q = sm.qs[p][kk][i];
            const float s_next = sm.sv[p][kk + 1];
            float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
            for (int cc = 0; cc < CPT; cc += 2) {
              const float2 l0 = make_float2(fmaf(s_cur, Rr[cc].x, Nr[cc].x), fmaf(s_cur, Rr[cc].y, Nr[cc].y));
              const float2 l1 = make_float2(fmaf(s_cur, Rr[cc + 1].x, Nr[cc + 1].x),
                                            fmaf(s_cur, Rr[cc + 1].y, Nr[cc + 1].y));
              cmac(a0, l0, xv[cc]);
              cmac(a1, l1, xv[cc + 1]);
            }
            float2 xp = make_float2(a0.x + a1.x, a0.y + a1.y);
            xp = group_sum_fast<NQ>(xp);
            const float2 xn = cmul(q, xp);
            sts_if(st2_on, st2 + kk * DP, (jq == 0) ? xn : xp);
            s_cur = s_next;
            chain_bar<NTC>();
          };
          if (len == CH) {
#pragma unroll 2
            for (int kk = 0; kk < CH; ++kk) step(kk);
          } else {
            for (int kk = 0; kk < len; ++kk) step(kk);
          }
          // chunk boundary: the un-scaled x_{k0+len} stays in ring p (the filler applies the scale
          // when it pulls the chunk); the scaled state goes straight into the next ring's slot 0
          float n2 = 0.f;
          for (int r = lane; r < DP; r += 32) n2 += cabs2(sm.xs[p][len][r]);
          n2 = warp_sum_f(n2);
          const float sc = rsqrtf(fmaxf(n2, 1e-12f));   // clamp of model.py:331-333
          if (t == 0) {
            sm.scv[p][0] = sc;
            if (scales) scales[(size_t)b * nchunks + c] = sc;
            mbar_arrive_remote(rfull + 8 * p);                 // ring p complete: tell the filler CTA
          }
          // ring p^1 is free once the filler has pulled chunk c-1 out of it
          mbar_wait(&sm.empty_bar[p ^ 1], (((c + 1) >> 1) & 1) ^ 1);
          if (t < DP) {
            const float2 v = sm.xs[p][len][t];
            sm.xs[p ^ 1][0][t] = make_float2(v.x * sc, v.y * sc);
          }
        }
      }
      __syncthreads();   // loader hand-over of chunk c+1's q_k / s_k
    }
    cluster_arrive_release();   // do not exit while the filler may still read this CTA's rings
    cluster_wait_acquire();
  } else {
    // ===================================== FILLER CTA =========================================
    const int tr = t;
    const bool work = tr < 2 * NTC;            // two work groups; the extra warp only copies and takes part in the barriers
    const int wg = tr / NTC;                   // work group: steps wg, wg + 2, ... of a chunk
    const int tl = tr % NTC;
    const int i = tl / NQ, jq = tl % NQ;
    float2 Sr[CPT];
    if (work) load_slice<DP, NQ>(Sr, matS, i, jq);
    double lossacc = 0.0;
    const unsigned rempty = dsmem_addr(&sm.empty_bar[0], 0);
    __syncthreads();
    for (int cc = 0; cc < nchunks; ++cc) {
      {
        const int p = cc & 1, len = chunk_len(cc), k0 = cc * CH;
        // waveform of chunk cc (for inc_k)
        for (int idx = t; idx <= len; idx += blockDim.x) cp_async4(&sm.wav[0][idx], xb + k0 + idx);
        cp_async_commit();
        // wait until the chain CTA has finished chunk cc, then pull it out of its rings
        mbar_wait(&sm.full_bar[p], (cc >> 1) & 1);
        {
          const unsigned rx = dsmem_addr(&sm.xs[p][0][0], 0);
          const unsigned rp = dsmem_addr(&sm.xps[p][0][0], 0);
          float4* lx = reinterpret_cast<float4*>(&sm.xs[0][0][0]);
          float4* lp = reinterpret_cast<float4*>(&sm.xps[0][0][0]);
          const float scp = ld_dsmem_f4(dsmem_addr(&sm.scv[p][0], 0)).x;   // c_k of this chunk
          for (int idx = t; idx < (len + 1) * DP / 2; idx += blockDim.x) {
            float4 v = ld_dsmem_f4(rx + 16 * idx);
            if (idx >= len * DP / 2) {     // row `len` = x_{k0+len}: stored un-scaled by the chain
              v.x *= scp;
              v.y *= scp;
              v.z *= scp;
              v.w *= scp;
            }
            lx[idx] = v;
          }
          for (int idx = t; idx < len * DP / 2; idx += blockDim.x) lp[idx] = ld_dsmem_f4(rp + 16 * idx);
        }
        cp_async_wait<0>();
        __syncthreads();
        if (t == 0) mbar_arrive_remote(rempty + 8 * p);       // ring p may be overwritten
        if (t < len) sm.incv[0][t] = sm.wav[0][t + 1] - sm.wav[0][t];
#ifdef AMPS_EXP_FWD_NOFILL
        if (len > 100000)
#endif
        if (work) {
          for (int kk = wg; kk < len; kk += 2) {
            const float2 part = matvec1<DP, NQ>(Sr, sm.xps[0][kk], jq);
            const float2 xpi = sm.xps[0][kk][i];
            sm.spp[kk][jq][i] = part;
            sm.es[kk][tl] = fmaf(xpi.x, part.x, xpi.y * part.y);
          }
        }
        __syncthreads();
        if (tr < NTC) {
          const int kk = tr / G, g = tr % G;
          float en = 0.f, nu2 = 0.f;
          if (kk < len) {
#pragma unroll 8
            for (int r = 0; r < NTC / G; ++r) en += sm.es[kk][g + G * r];
            for (int r = g; r < DP; r += G) nu2 += cabs2(sm.xs[0][kk][r]);
          }
#pragma unroll
          for (int m = 1; m < G && m < 32; m <<= 1) {
            en += __shfl_xor_sync(0xffffffffu, en, m);
            nu2 += __shfl_xor_sync(0xffffffffu, nu2, m);
          }
          if (g == 0 && kk < len) {
            const float E = en / fmaxf(nu2, 1e-12f);                                  // model.py:324-325 on x'
            const float z = (E * sm.incv[0][kk]) / A;                  // model.py:294
            lossacc -= (double)log1pf(z);
            if (evout) evout[(size_t)b * T + k0 + kk] = make_float2(E, nu2);   // for the adjoint sweep
          }
        }
        if (seg.ckpt && cc % seg.ck_chunks == 0 && t < DP)    // state checkpoint: x at the start of chunk cc
          seg.ckpt[(size_t)b * seg.ck_stride + (size_t)(cc / seg.ck_chunks) * DP + t] = sm.xs[0][0][t];
#ifdef AMPS_EXP_FWD_NOFILL
        if (len > 100000)
#endif
        if (traj) {
          const float4* src = reinterpret_cast<const float4*>(&sm.xs[0][1][0]);
          float4* dst = reinterpret_cast<float4*>(traj + ((size_t)b * T + k0 + 1) * DP);
          for (int idx = t; idx < len * DP / 2; idx += blockDim.x) dst[idx] = src[idx];
        }
#ifdef AMPS_EXP_FWD_NOFILL
        if (len > 100000)
#endif
        if (sptraj) {   // S x'_k for the adjoint sweep (saves it the mat-vec)
          float2* dst = sptraj + ((size_t)b * T + k0) * DP;
          for (int idx = t; idx < len * DP; idx += blockDim.x) {
            const int kk = idx / DP, r = idx % DP;
            float2 sp = sm.spp[kk][0][r];
#pragma unroll
            for (int j = 1; j < NQ; ++j) {
              sp.x += sm.spp[kk][j][r].x;
              sp.y += sm.spp[kk][j][r].y;
            }
            dst[idx] = sp;
          }
        }
      }
      __syncthreads();
    }
    cluster_arrive_release();
    cluster_wait_acquire();
    lossacc = warp_sum_d(lossacc);
    if (lane == 0) sm.lred[t >> 5] = lossacc;
    __syncthreads();
    if (t == 0) {
      double tot = 0.0;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) tot += sm.lred[wv];
      loss[b] = (float)tot;
      if (lossd) lossd[b] = tot;
    }
  }
}

// -------------------------------------------------------------------------------------------
// backward, cluster variant.  block = 3*NTC threads in both CTAs (NTC = DP*NQ).
//   rank 0 CHAIN CTA : threads 0..127 run the adjoint recursion; it only ever touches its own
//                      shared memory (packed inputs pushed by the filler, mu ring).
//   rank 1 FILLER CTA: threads 0..127 "prep" group  -- cp.async of trajectory / q / waveform, x', S x',
//                      alpha, beta, packed rows pushed into the chain CTA (st.shared::cluster);
//                      two "tiles" groups of NTC threads -- pull the finished chunk's mu ring
//                      (ld.shared::cluster) and accumulate the rank-1 gradient tiles, each over
//                      half of the chunk's steps (their accumulators are summed at the end).
// Slot c (c = nchunks .. -1): chain runs chunk c; prep prepares chunk c-1 and loads chunk c-2;
// tiles accumulate chunk c+1.  Hand-off by mbarriers in each other's shared memory (remote
// mbarrier.arrive.release.cluster, local try_wait), two buffers by chunk parity:
//   done_bar[p] (filler CTA): the chain has finished the chunk in buffer p -- its mu ring is complete and
//                             its pushed inputs are dead; the filler then pulls the ring and pushes the
//                             inputs of the chunk two further down into the same buffer
//   in_full[p]  (chain CTA) : both of those have happened: the chain may run the next chunk in buffer p
// The chain never waits on a cluster-wide barrier (one barrier.cluster per chunk cost ~1000 cycles).
// -------------------------------------------------------------------------------------------
template <int DP, int NQ>
struct alignas(16) BwdClChainSmem {   // what the CHAIN CTA keeps (and the filler addresses remotely)
  float4 cina[2][CH][DP];      // { c_k q_k , alpha_k (S x'_k)_i }
  float4 cinb[2][CH][DP];      // { beta_k x_k,i , dtm_k x_k,i }
  float2 mus[2][CH][DP];       // adjoint of x'_k
  float svc[2][CH];            // s_k for the chain
  unsigned long long in_full[2];   // buffer p holds the pushed inputs of a chunk AND its mu ring has been pulled
};
template <int DP, int NQ>
struct alignas(16) BwdClFillSmem {    // FILLER CTA
  static constexpr int NTC = DP * NQ, G = NTC / CH, ES = NTC + (G < 32 ? G : 0);
  float2 xs[4][CH + 1][DP];
  float2 qs[2][CH][DP];
  float2 xps[3][CH][DP];
  float2 mul[CH][DP];          // local copy of the pulled mu ring
  float2 spl[2][CH][DP];       // S x'_k stored by the forward
  float2 evl[2][CH];           // (E_k, |x_k|^2) stored by the forward
  float wav[2][CH + 4];
  float tt[2][CH + 4];
  float scs[2][4];
  float sv[3][CH], alphas[3][CH];
  float incv[CH], betas[CH], dtm[CH];
  double lred[32];
  unsigned long long done_bar[2];  // the chain CTA has finished the chunk in buffer p
};
template <int DP, int NQ>
constexpr size_t bwd_cl_smem_bytes() {
  return sizeof(BwdClChainSmem<DP, NQ>) > sizeof(BwdClFillSmem<DP, NQ>) ? sizeof(BwdClChainSmem<DP, NQ>)
                                                                        : sizeof(BwdClFillSmem<DP, NQ>);
}

// launch bound 512 (the kernel runs 3*DP*NQ <= 384 threads): caps the chain at 128 registers, so that a
// 160-thread forward-replay CTA still fits next to a 384-thread adjoint CTA on one SM (checkpointed backward)
#ifndef AMPS_BWD_CL_MAXT
#define AMPS_BWD_CL_MAXT 384
#endif
#ifndef AMPS_BWD_CL_MINB
#define AMPS_BWD_CL_MINB 0
#endif
#ifndef AMPS_BWD_PREFETCH
#define AMPS_BWD_PREFETCH 1
#endif
template <int DP, int NQ>
#if AMPS_BWD_CL_MINB
__global__ void __launch_bounds__(AMPS_BWD_CL_MAXT, 1)
#else
__global__ void __launch_bounds__(AMPS_BWD_CL_MAXT)
#endif
    psi_bwd_cl_kernel(const float2* __restrict__ matN, const float2* __restrict__ matRH,
                      const float2* __restrict__ matS, const float2* __restrict__ qtab,
                      const float* __restrict__ ttab, const float* __restrict__ x, int T, AVal A_,
                      const float* __restrict__ w, const float2* __restrict__ traj,
                      const float* __restrict__ scales, int nchunks, float2* __restrict__ Gout,
                      float* __restrict__ gfout, float2* __restrict__ lam0out,
                      double* __restrict__ gAdir, const float2* __restrict__ sptraj,
                      const float2* __restrict__ evin, SegBwd seg) {
  const float A = a_get(A_);
  using M = Map<DP, NQ>;
  constexpr int NTC = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int NP = M::NP;
  constexpr int LV = (NQ == 4) ? 2 : 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // both views alias the same buffer; each CTA uses one, the filler addresses the chain's remotely
  BwdClChainSmem<DP, NQ>& cs = *reinterpret_cast<BwdClChainSmem<DP, NQ>*>(smem_raw);
  BwdClFillSmem<DP, NQ>& sm = *reinterpret_cast<BwdClFillSmem<DP, NQ>*>(smem_raw);

  const int t = threadIdx.x;
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x >> 1;
  const int nsteps = T - 1;
  const int lane = t & 31;
  const int grp = t / NTC;                 // chain CTA: 0 = chain; filler CTA: 0 = prep, 1..2 = tiles
  const int tr = t - grp * NTC;
  const int i = tr / NQ, jq = tr % NQ;
  auto chunk_len = [&](int c) { return min(CH, nsteps - c * CH); };
  // n-th use (descending chunks) of the parity-p barriers by chunk c: phase parity of that use
  auto use_parity = [&](int c) { return (unsigned)((((nchunks - 1 - c) >> 1)) & 1); };

  if (t == 0) {
    if (rank == 0) {
      mbar_init(&cs.in_full[0], 2);   // one arrival from the prep group (inputs pushed), one from the tiles
      mbar_init(&cs.in_full[1], 2);   // groups (ring pulled)
    } else {
      mbar_init(&sm.done_bar[0], 1);
      mbar_init(&sm.done_bar[1], 1);
    }
    mbar_fence_init_cluster();
  }
  __syncthreads();
  cluster_arrive_release();   // both CTAs' barriers are initialised before any remote arrive
  cluster_wait_acquire();

  if (rank == 0) {
    // ===================================== CHAIN CTA ==========================================
    float2 Nr[CPT], Hr[CPT];
    float2 lam = make_float2(0.f, 0.f);
    float gf = 0.f;
    if (grp == 0) {
      load_slice<DP, NQ>(Nr, matN, i, jq);
      load_slice<DP, NQ>(Hr, matRH, i, jq);
      if (seg.lam_end) lam = seg.lam_end[(size_t)b * DP + i];
      if (seg.accumulate) gf = gfout[(size_t)b * DP + i];
    }
    const bool mu_on = jq == 0;
    const unsigned rdone = dsmem_addr(&sm.done_bar[0], 1);
    if (grp == 0) for (int c = nchunks - 1; c >= 0; --c) {
      {
        const int ca = c & 1, len = chunk_len(c);
        mbar_wait(&cs.in_full[ca], use_parity(c));   // inputs of chunk c pushed, ring ca pulled
        // shared-window addresses of this chunk's chain inputs, computed once (generic pointers into
        // shared memory make the compiler re-read %cluster_ctaid inside the step loop)
        const unsigned mu_a = smem_addr_pinned(&cs.mus[ca][0][i]);
        const unsigned cina_a = smem_addr_pinned(&cs.cina[ca][0][i]);
        const unsigned cinb_a = smem_addr_pinned(&cs.cinb[ca][0][i]);
        const unsigned svc_a = smem_addr_pinned(&cs.svc[ca][0]);
        constexpr unsigned ROW4 = DP * sizeof(float4), ROW2 = DP * sizeof(float2);
        {
          const float4 a4 = lds128v(cina_a + (unsigned)(len - 1) * ROW4);
          float2 mu = cmul_ca(make_float2(a4.x, a4.y), lam);
          mu.x += a4.z;
          mu.y += a4.w;
          sts64a_if(mu_on, mu_a + (unsigned)(len - 1) * ROW2, mu);
        }
        chain_bar<NTC>();
        const unsigned mus_addr = smem_addr_pinned(&cs.mus[ca][0][2 * jq]);
        // the filler's inputs of a step (s_k, {beta x_k, dt x_k}, {c q, alpha S x'}_{k-1}) are complete for the
        // whole chunk before the chain enters it: they are fetched ONE STEP AHEAD, behind the mu loads, so
        // that after the barrier only the mu ring itself is waited for and L_k = N + s_k R^dag can be formed
        // while those loads are in flight
        float s = lds32a(svc_a + (unsigned)(len - 1) * (unsigned)sizeof(float));
        float4 b4 = lds128v(cinb_a + (unsigned)(len - 1) * ROW4);
        float4 a4 = lds128v(cina_a + (unsigned)(len > 1 ? len - 2 : 0) * ROW4);
        auto step = [&](int kk) {
          float2 mv[CPT];
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            const float4 v = lds128v(mus_addr + (unsigned)((kk * DP + 2 * NQ * m) * sizeof(float2)));
            mv[2 * m] = make_float2(v.x, v.y);
            mv[2 * m + 1] = make_float2(v.z, v.w);
          }
          const unsigned km = (unsigned)(kk > 0 ? kk - 1 : 0);
          const unsigned km2 = (unsigned)(kk > 1 ? kk - 2 : 0);
#if AMPS_BWD_PREFETCH
          const float s_n = lds32a(svc_a + km * (unsigned)sizeof(float));
          const float4 b4_n = lds128v(cinb_a + km * ROW4);
          const float4 a4_n = lds128v(cina_a + km2 * ROW4);
#else
          (void)km2;
          s = lds32a(svc_a + (unsigned)kk * (unsigned)sizeof(float));
          b4 = lds128v(cinb_a + (unsigned)kk * ROW4);
          a4 = lds128v(cina_a + km * ROW4);
#endif
          float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
          for (int cc = 0; cc < CPT; cc += 2) {
            const float2 l0 = make_float2(fmaf(s, Hr[cc].x, Nr[cc].x), fmaf(s, Hr[cc].y, Nr[cc].y));
            const float2 l1 = make_float2(fmaf(s, Hr[cc + 1].x, Nr[cc + 1].x),
                                          fmaf(s, Hr[cc + 1].y, Nr[cc + 1].y));
            cmac(a0, l0, mv[cc]);
            cmac(a1, l1, mv[cc + 1]);
          }
          float2 lp = make_float2(a0.x + a1.x, a0.y + a1.y);
          lp = group_sum_fast<NQ>(lp);
          lam.x = lp.x + b4.x;
          lam.y = lp.y + b4.y;
          gf = fmaf(lam.x, b4.w, fmaf(-lam.y, b4.z, gf));
          float2 mu = cmul_ca(make_float2(a4.x, a4.y), lam);
          mu.x += a4.z;
          mu.y += a4.w;
          sts64a_if(mu_on && kk > 0, mu_a + km * ROW2, mu);
#if AMPS_BWD_PREFETCH
          s = s_n;
          b4 = b4_n;
          a4 = a4_n;
#endif
          chain_bar<NTC>();
        };
        if (len == CH) {
#pragma unroll 2
          for (int kk = CH - 1; kk >= 0; --kk) step(kk);
        } else {
          for (int kk = len - 1; kk >= 0; --kk) step(kk);
        }
        // every chain thread is past the last step's barrier: ring ca is complete, its inputs are dead
        if (t == 0) mbar_arrive_remote(rdone + 8 * ca);
      }
    }
    if (grp == 0 && jq == 0) {
      gfout[(size_t)b * DP + i] = gf;
      lam0out[(size_t)b * DP + i] = lam;
    }
    cluster_arrive_release();   // do not exit while the filler may still read this CTA's rings
    cluster_wait_acquire();
  } else {
    // ===================================== FILLER CTA =========================================
    const float* xb = x + (size_t)b * seg.xstride;
    const float2* trb = traj + (size_t)b * T * DP;
    const float wb = w[b];
    float2 GR[CPT], GN[CPT], GE[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) GR[c] = GN[c] = GE[c] = make_float2(0.f, 0.f);
    double gAacc = 0.0;
    const bool tpv = seg.tprev_valid != 0;

    auto issue_loads = [&](int c) {          // prep group
      const int k0 = c * CH, len = chunk_len(c);
      const float2* xsrc = trb + (size_t)k0 * DP;
      float2* xdst = &sm.xs[c & 3][0][0];
      for (int idx = tr; idx < (len + 1) * DP / 2; idx += NTC) cp_async16(xdst + 2 * idx, xsrc + 2 * idx);
      const float2* qsrc = qtab + (size_t)k0 * DP;
      float2* qdst = &sm.qs[c & 1][0][0];
      for (int idx = tr; idx < len * DP / 2; idx += NTC) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
      for (int idx = tr; idx <= len; idx += NTC) {
        cp_async4(&sm.wav[c & 1][idx], xb + k0 + idx);
        cp_async4(&sm.tt[c & 1][idx], ttab + ((k0 + idx > 0 || tpv) ? k0 + idx - 1 : 0));
      }
      if (tr == 0) cp_async4(&sm.scs[c & 1][0], scales + (size_t)b * nchunks + c);
      const float2* ssrc = sptraj + ((size_t)b * T + k0) * DP;
      float2* sdst = &sm.spl[c & 1][0][0];
      for (int idx = tr; idx < len * DP / 2; idx += NTC) cp_async16(sdst + 2 * idx, ssrc + 2 * idx);
      const float2* esrc = evin + (size_t)b * T + k0;   // 8-byte elements: two 4-byte copies each
      for (int idx = tr; idx < 2 * len; idx += NTC)
        cp_async4(reinterpret_cast<float*>(&sm.evl[c & 1][0]) + idx, reinterpret_cast<const float*>(esrc) + idx);
    };

    auto tiles_pull = [&](int c) {           // tiles group: the chunk's mu ring out of the chain CTA
      const int len = chunk_len(c);
      const unsigned rm = dsmem_addr(&cs.mus[c & 1][0][0], 0);
      float4* lm = reinterpret_cast<float4*>(&sm.mul[0][0]);
      for (int idx = t - NTC; idx < len * DP / 2; idx += 2 * NTC) lm[idx] = ld_dsmem_f4(rm + 16 * idx);
    };
    auto tiles_chunk = [&](int c) {          // tiles group (after the pull's barrier: sm.mul is complete)
      const int len = chunk_len(c);
      const float2(*xsb)[DP] = sm.xs[c & 3];
      const float2(*xpb)[DP] = sm.xps[c % 3];
#ifdef AMPS_EXPERIMENT_NO_TILES
      if (len > 100000)
#endif
      for (int kk = grp - 1; kk < len; kk += 2) {   // the two tile groups interleave the steps
        const float2 mui = sm.mul[kk][i];
        const float2 xpi = xpb[kk][i];
        const float s = sm.sv[c % 3][kk];
        const float al = sm.alphas[c % 3][kk];
        const float2 u1 = make_float2(s * mui.x, s * mui.y);
        const float2 u3 = make_float2(al * xpi.x, al * xpi.y);
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(&xsb[kk][2 * NQ * m + 2 * jq]);
          const float4 pv = *reinterpret_cast<const float4*>(&xpb[kk][2 * NQ * m + 2 * jq]);
          const float2 x0 = make_float2(xv.x, xv.y), x1 = make_float2(xv.z, xv.w);
          const float2 p0 = make_float2(pv.x, pv.y), p1 = make_float2(pv.z, pv.w);
          cmac_cx(GR[2 * m], u1, x0);
          cmac_cx(GR[2 * m + 1], u1, x1);
          cmac_cx(GN[2 * m], mui, x0);
          cmac_cx(GN[2 * m + 1], mui, x1);
          cmac_cx(GE[2 * m], u3, p0);
          cmac_cx(GE[2 * m + 1], u3, p1);
        }
      }
    };

    auto prep_chunk = [&](int c) {           // prep group: elementwise only
      const int len = chunk_len(c), k0 = c * CH;
      const int lx = c & 3, lq = c & 1, lp3 = c % 3;
      const float sc = sm.scs[lq][0];
      const float inv_sc = 1.0f / sc;
      const unsigned ra = dsmem_addr(&cs.cina[lq][0][0], 0);
      const unsigned rb = dsmem_addr(&cs.cinb[lq][0][0], 0);
      const unsigned rs = dsmem_addr(&cs.svc[lq][0], 0);
      if (tr < len) {
        const float inc = sm.wav[lq][tr + 1] - sm.wav[lq][tr];
        const float s = inc / A;
        sm.sv[lp3][tr] = s;
        st_dsmem_f1(rs + 4 * tr, s);
        sm.dtm[tr] = (k0 + tr > 0 || tpv) ? sm.tt[lq][tr + 1] - sm.tt[lq][tr] : 0.f;
        const float2 ev = sm.evl[lq][tr];          // (E_k, |x_k|^2) from the forward
        const float E = ev.x, nu2 = ev.y;
        const float arg = 1.0f + (E * inc) / A;
        const float gE = wb * (-s / arg);
        const float alpha = 2.0f * gE / fmaxf(nu2, 1e-12f);
        sm.alphas[lp3][tr] = alpha;
        sm.betas[tr] = -alpha * E;
        gAacc += (double)wb * (double)E * (double)inc / ((double)A * (double)A * (double)arg);
      }
      bar_named(2, NTC);
      // x'_k = conj(q_k) x_{k+1} / c_k (for the tiles) and the packed per-row chain inputs,
      // pushed straight into the chain CTA's shared memory
      for (int idx = tr; idx < len * DP; idx += NTC) {
        const int kk = idx / DP, r = idx % DP;
        float2 q = sm.qs[lq][kk][r];
        const float2 xk = sm.xs[lx][kk][r];
        float2 xp = cmul_ca(q, sm.xs[lx][kk + 1][r]);
        if (kk == len - 1) {
          xp.x *= inv_sc;
          xp.y *= inv_sc;
          q.x *= sc;
          q.y *= sc;
        }
        sm.xps[lp3][kk][r] = xp;
        const float al = sm.alphas[lp3][kk], be = sm.betas[kk], dt = sm.dtm[kk];
        const float2 sp = sm.spl[lq][kk][r];
        st_dsmem_f4(ra + 16 * idx, make_float4(q.x, q.y, al * sp.x, al * sp.y));
        st_dsmem_f4(rb + 16 * idx, make_float4(be * xk.x, be * xk.y, dt * xk.x, dt * xk.y));
      }
    };

    if (grp == 0) {
      if (nchunks > 0) issue_loads(nchunks - 1);
      cp_async_commit();
    }
    const unsigned rfull = dsmem_addr(&cs.in_full[0], 0);
    for (int c = nchunks; c >= -1; --c) {
      const bool have_up = c + 1 >= 0 && c + 1 < nchunks;   // chunk c+1 exists: its ring is to be pulled
      // the chain has finished chunk c+1: ring (c+1)&1 is complete and buffer (c-1)&1 == (c+1)&1 is free
      if (have_up) mbar_wait(&sm.done_bar[(c + 1) & 1], use_parity(c + 1));
      if (grp == 0) {
        if (c - 2 >= 0) issue_loads(c - 2);
        cp_async_commit();
        cp_async_wait<1>();
        bar_named(2, NTC);
        if (c - 1 >= 0) {
          prep_chunk(c - 1);
          bar_named(2, NTC);      // every prep thread's pushes are issued
          if (tr == 0) mbar_arrive_remote(rfull + 8 * ((c - 1) & 1));
        }
      } else {
        if (have_up) tiles_pull(c + 1);
        bar_named(3, 2 * NTC);    // ring pulled (sm.mul complete)
        if (t == NTC && c - 1 >= 0) mbar_arrive_remote(rfull + 8 * ((c - 1) & 1));
        if (have_up) tiles_chunk(c + 1);
      }
      __syncthreads();          // tiles done with xps / xs / mul before the next slot's prep rewrites them
    }
    cluster_arrive_release();
    cluster_wait_acquire();
    if (grp == 0) {
      cp_async_wait<0>();
      gAacc = warp_sum_d(gAacc);
      if (lane == 0) sm.lred[tr >> 5] = gAacc;
      bar_named(2, NTC);
      if (tr == 0) {
        double tot = seg.accumulate ? gAdir[b] : 0.0;
        for (int wv = 0; wv < NTC / 32; ++wv) tot += sm.lred[wv];
        gAdir[b] = tot;
      }
    } else {
      // sum the two tile groups' accumulators through shared memory (xs is free now)
      float2* scratch = &sm.xs[0][0][0];
      if (grp == 2) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          scratch[(0 * CPT + c) * NTC + tr] = GR[c];
          scratch[(1 * CPT + c) * NTC + tr] = GN[c];
          scratch[(2 * CPT + c) * NTC + tr] = GE[c];
        }
      }
      bar_named(3, 2 * NTC);
      if (grp == 1) {
        float2* Gb = Gout + (size_t)b * 3 * DP * DP;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const int col = M::col(c, jq);
          float2 r = scratch[(0 * CPT + c) * NTC + tr], n = scratch[(1 * CPT + c) * NTC + tr],
                 e = scratch[(2 * CPT + c) * NTC + tr];
          r = make_float2(GR[c].x + r.x, GR[c].y + r.y);
          n = make_float2(GN[c].x + n.x, GN[c].y + n.y);
          e = make_float2(GE[c].x + e.x, GE[c].y + e.y);
          if (seg.accumulate) {   // continue the sums of the later time windows
            const float2 r0 = Gb[0 * DP * DP + i * DP + col], n0 = Gb[1 * DP * DP + i * DP + col],
                         e0 = Gb[2 * DP * DP + i * DP + col];
            r = make_float2(r.x + r0.x, r.y + r0.y);
            n = make_float2(n.x + n0.x, n.y + n0.y);
            e = make_float2(e.x + e0.x, e.y + e0.y);
          }
          Gb[0 * DP * DP + i * DP + col] = r;
          Gb[1 * DP * DP + i * DP + col] = n;
          Gb[2 * DP * DP + i * DP + col] = e;
        }
      }
    }
  }
}

}  // namespace amps
