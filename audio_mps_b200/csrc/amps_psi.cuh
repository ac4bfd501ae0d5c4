// PsiCMPS scan kernels (sm_100a): sequential persistent forward-loss kernel, adjoint backward,
// sampler.  One CTA owns one clip for the whole clip.
//
// Formulation (validated against the op-for-op oracle, see DESIGN.md "Chain form"):
// in the interaction frame x_k = psi_k * conj(p_k) the reference step (model.py:276-334) is
//     x'_k    = L_k x_k,   L_k = N + s_k R,  N = I - (delta_t sigma^2 / 2) R^dag R,  s_k = inc_k / A
//     E_k     = x'_k^dag (R + R^dag) x'_k / |x_k|^2
//     loss   += -log(1 + (E_k inc_k) / A)
//     x_{k+1} = q_k * x'_k (* c),             q_k = p_k conj(p_{k+1})
// with p_k = exp(i fl32(f t_k)) and t_k the float32 running sum.  The state is carried
// UN-normalised (the loss is scale invariant) and rescaled by c once per chunk of CHK steps.
//
// Warp specialisation.  The recursion is a dependent chain
//     LDS state -> FFMA (mat-vec with L_k) -> 2..3 shuffle levels -> STS -> bar.sync
// whose measured B200 latencies (profiles/micro: shfl+add 29 cycles, STS+bar+LDS 50 cycles, a
// dependent FFMA stream 1.4-1.6 cycles/instruction with one warp per SM sub-partition) leave most
// issue slots empty.  Each CTA therefore runs two sets of warps on the same sub-partitions:
//   * CHAIN warps (high warp ids): nothing but the recursion; one named barrier per step.
//   * FILLER warps: everything that does not feed the next state -- cp.async prefetch of the
//     waveform / phase table / trajectory, s_k, the expectation mat-vec S x', the per-step
//     scalars (E_k, |x_k|^2, log1p, alpha_k, beta_k), the rank-1 gradient tiles, the trajectory
//     flush -- one chunk behind (forward) or one chunk ahead/behind (backward), with no per-step
//     synchronisation.  The hardware scheduler interleaves the two instruction streams.
// The two sets meet at one __syncthreads per chunk.
#pragma once
#include <type_traits>

#include "amps_common.cuh"

namespace amps {

// steps per chunk for a padded bond dimension (shared-memory budget of the backward kernel)
__host__ __device__ constexpr int chunk_steps(int) { return CH; }

__device__ __forceinline__ void bar_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// chain barrier of NT threads: a chain of ONE warp (DP = 8: 32 chain threads) only needs warp-level
// ordering of its shared-memory traffic, not a bar.sync round trip
template <int NT>
__device__ __forceinline__ void chain_bar() {
  if (NT == 32) __syncwarp();
  else bar_named(1, NT);
}

// -------------------------------------------------------------------------------------------
// shared-memory layouts
// -------------------------------------------------------------------------------------------
template <int DP, int NQ>
struct alignas(16) FwdSmem {
  static constexpr int CHK = chunk_steps(DP);
  static constexpr int NTC = DP * NQ, G = NTC / CHK, ES = NTC + (G < 32 ? G : 0);
  float2 xs[2][CHK + 1][DP];   // x_{k0+kk}, ring by chunk parity
  float2 xps[2][CHK][DP];      // x'_{k0+kk}
  float2 qs[2][CHK][DP];       // q_k
  float es[CHK][ES];           // per-thread partial of Re(x'^dag S x')
  float2 spp[CHK][NQ][DP];     // per-lane partials of S x'_k (summed when flushed for the backward)
  float wav[2][CHK + 4];       // waveform samples k0..k0+len
  float sv[2][CHK + 4];        // s_k
  float incv[2][CHK];          // inc_k
  double lred[32];
};

template <int DP, int NQ>
struct alignas(16) BwdSmem {
  static constexpr int CHK = chunk_steps(DP);
  float2 xs[4][CHK + 1][DP];   // trajectory chunk (chunk % 4): landing / prep / chain / tiles
  float2 qs[2][CHK][DP];       // q_k (chunk & 1): landing / prep
  float2 xps[3][CHK][DP];      // reconstructed x'_k (chunk % 3): prep / - / tiles
  float4 cina[2][CHK][DP];     // chain inputs (chunk & 1): { c_k q_k , alpha_k (S x'_k)_i }
  float4 cinb[2][CHK][DP];     //                           { beta_k x_k,i , dtm_k x_k,i }
  float2 mus[2][CHK][DP];      // adjoint of x'_k (chunk & 1): chain writes, tiles read
  float2 spl[2][CHK][DP];      // S x'_k stored by the forward (chunk & 1): landing / prep
  float2 evl[2][CHK];          // (E_k, |x_k|^2) stored by the forward
  float wav[2][CHK + 4];
  float tt[2][CHK + 4];        // tt[.][0] = t_{k0-1}, tt[.][1+kk] = t_{k0+kk}
  float scs[2][4];
  float sv[3][CHK], alphas[3][CHK];
  float incv[CHK], betas[CHK], dtm[CHK];
  double lred[32];
};

template <int DP>
struct alignas(16) SampleSmem {
  float2 xs[2][DP];
  float2 qs[2][CH][DP];
  float nz[2][CH];
  float outs[CH];
  float wred[2][32][2];
};

// -------------------------------------------------------------------------------------------
// K1: forward per-clip loss (model.py:257-267, 276-282, 293-334)
// -------------------------------------------------------------------------------------------
template <int DP, int NQ>
__global__ void __launch_bounds__(2 * DP * NQ, 1)
    psi_fwd_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                   const float2* __restrict__ matS, const float2* __restrict__ qtab,
                   const float2* __restrict__ psi0p, const float* __restrict__ x, int T, AVal A_,
                   float* __restrict__ loss, double* __restrict__ lossd,
                   float2* __restrict__ traj, float* __restrict__ scales, int nchunks,
                   float2* __restrict__ sptraj, float2* __restrict__ evout, SegFwd seg) {
  const float A = a_get(A_);
  using M = Map<DP, NQ>;
  using Sm = FwdSmem<DP, NQ>;
  constexpr int NTC = M::NT;       // threads per role
  constexpr int CPT = M::CPT;
  constexpr int CHK = Sm::CHK;
  constexpr int G = Sm::G;         // filler threads per step in the scalar phase
  constexpr int LV = (NQ == 4) ? 2 : 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm& sm = *reinterpret_cast<Sm*>(smem_raw);

  const int t = threadIdx.x;
  const bool is_chain = t >= NTC;          // high warp ids: the scheduler favours them
  const int tr = is_chain ? t - NTC : t;   // thread index within the role
  const int i = tr / NQ, jq = tr % NQ, lane = t & 31;
  const int b = blockIdx.x;
  const int nsteps = T - 1;
  const float* xb = x + (size_t)b * seg.xstride;
  auto chunk_len = [&](int c) { return min(CHK, nsteps - c * CHK); };

  // ---- prologue (all threads): inputs of chunk 0, start state ------------------------------
  if (t < DP) {
    const float2 p = seg.x0 ? seg.x0[(size_t)b * seg.x0_stride + t] : psi0p[t];
    sm.xs[0][0][t] = p;
    if (traj) traj[(size_t)b * T * DP + t] = p;
  }
  auto issue_loads = [&](int c, int nthr, int tid) {
    const int buf = c & 1, k0 = c * CHK, len = chunk_len(c);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[buf][0][0];
    for (int idx = tid; idx < len * DP / 2; idx += nthr) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = tid; idx <= len; idx += nthr) cp_async4(&sm.wav[buf][idx], xb + k0 + idx);
  };
  auto compute_s = [&](int c, int tid) {
    const int buf = c & 1, len = chunk_len(c);
    if (tid < len) {
      const float inc = sm.wav[buf][tid + 1] - sm.wav[buf][tid];   // model.py:263
      sm.incv[buf][tid] = inc;
      sm.sv[buf][tid] = inc / A;                                    // model.py:303
    }
  };
  if (nchunks > 0) {
    issue_loads(0, 2 * NTC, t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    compute_s(0, t);
  }
  __syncthreads();

  if (is_chain) {
    // =================================== CHAIN WARPS ==========================================
    float2 Nr[CPT], Rr[CPT];
    load_slice<DP, NQ>(Nr, matN, i, jq);
    load_slice<DP, NQ>(Rr, matR, i, jq);
    float2 vstart = make_float2(0.f, 0.f);
    for (int c = 0; c <= nchunks; ++c) {
      if (c < nchunks) {
        const int p = c & 1, len = chunk_len(c);
        if (c > 0) {
          if (tr < DP) sm.xs[p][0][tr] = vstart;   // fillers are done with ring p (chunk c-2)
          bar_named(1, NTC);
        }
        // branch-free stores: lane jq==0 writes x_{k+1,i}, jq==1 writes x'_{k,i}
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
          for (int cc = 0; cc < CPT; cc += 2) {   // L_k slice formed in the shadow of the loads
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
          bar_named(1, NTC);
        };
        if (len == CHK) {
#pragma unroll 2
          for (int kk = 0; kk < CHK; ++kk) step(kk);
        } else {
          for (int kk = 0; kk < len; ++kk) step(kk);
        }
        // rescale by 1/|x_{k0+len}| (each chain warp computes the norm redundantly)
        float n2 = 0.f;
        for (int r = lane; r < DP; r += 32) n2 += cabs2(sm.xs[p][len][r]);
        n2 = warp_sum_f(n2);
        const float sc = rsqrtf(fmaxf(n2, 1e-12f));   // clamp of model.py:331-333
        bar_named(1, NTC);   // every chain warp has read the un-scaled state
        if (tr < DP) {
          float2 v = sm.xs[p][len][tr];
          v.x *= sc;
          v.y *= sc;
          sm.xs[p][len][tr] = v;   // the scaled state is what the trajectory / backward sees
          vstart = v;
        }
        if (tr == 0 && scales) scales[(size_t)b * nchunks + c] = sc;
      }
      __syncthreads();   // chunk hand-over
    }
  } else {
    // =================================== FILLER WARPS =========================================
    float2 Sr[CPT];
    load_slice<DP, NQ>(Sr, matS, i, jq);
    double lossacc = 0.0;
    for (int c = 0; c <= nchunks; ++c) {
      if (c + 1 < nchunks) issue_loads(c + 1, NTC, tr);
      cp_async_commit();
      if (c >= 1) {
        // ---- chunk c-1: expectation mat-vec for every step (no per-step synchronisation) ----
        const int cc = c - 1, p = cc & 1, len = chunk_len(cc), k0 = cc * CHK;
#ifndef AMPS_EXPERIMENT_NO_FILLER
        for (int kk = 0; kk < len; ++kk) {
          const float2 part = matvec1<DP, NQ>(Sr, sm.xps[p][kk], jq);
          const float2 xpi = sm.xps[p][kk][i];
          sm.spp[kk][jq][i] = part;
          sm.es[kk][tr] = fmaf(xpi.x, part.x, xpi.y * part.y);
        }
#endif
        bar_named(2, NTC);
        {  // per-step scalars: G threads per step
          const int kk = tr / G, g = tr % G;
          float en = 0.f, nu2 = 0.f;
          if (kk < len) {
#pragma unroll 8
            for (int r = 0; r < NTC / G; ++r) en += sm.es[kk][g + G * r];
            for (int r = g; r < DP; r += G) nu2 += cabs2(sm.xs[p][kk][r]);
          }
#pragma unroll
          for (int m = 1; m < G && m < 32; m <<= 1) {
            en += __shfl_xor_sync(0xffffffffu, en, m);
            nu2 += __shfl_xor_sync(0xffffffffu, nu2, m);
          }
          if (g == 0 && kk < len) {
            const float E = en / fmaxf(nu2, 1e-12f);                                  // model.py:324-325 on x'
            const float z = (E * sm.incv[p][kk]) / A;                  // model.py:294
            lossacc -= (double)log1pf(z);
            if (evout) evout[(size_t)b * T + k0 + kk] = make_float2(E, nu2);   // for the adjoint sweep
          }
        }
        if (seg.ckpt && cc % seg.ck_chunks == 0 && tr < DP)   // state checkpoint: x at the start of chunk cc
          seg.ckpt[(size_t)b * seg.ck_stride + (size_t)(cc / seg.ck_chunks) * DP + tr] = sm.xs[p][0][tr];
        if (traj) {   // flush x_{k0+1 .. k0+len}
          const float4* src = reinterpret_cast<const float4*>(&sm.xs[p][1][0]);
          float4* dst = reinterpret_cast<float4*>(traj + ((size_t)b * T + k0 + 1) * DP);
          for (int idx = tr; idx < len * DP / 2; idx += NTC) dst[idx] = src[idx];
        }
        if (sptraj) {   // S x'_k for the adjoint sweep (saves it the mat-vec)
          float2* dst = sptraj + ((size_t)b * T + k0) * DP;
          for (int idx = tr; idx < len * DP; idx += NTC) {
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
      cp_async_wait<0>();
      bar_named(2, NTC);   // chunk c+1's waveform has landed for every filler thread
      if (c + 1 < nchunks) compute_s(c + 1, tr);
      __syncthreads();     // chunk hand-over
    }
    // block reduction of the per-thread loss partials (filler warps only)
    lossacc = warp_sum_d(lossacc);
    if (lane == 0) sm.lred[tr >> 5] = lossacc;
    bar_named(2, NTC);
    if (tr == 0) {
      double tot = 0.0;
      for (int wv = 0; wv < NTC / 32; ++wv) tot += sm.lred[wv];
      loss[b] = (float)tot;
      if (lossd) lossd[b] = tot;
    }
  }
}

// -------------------------------------------------------------------------------------------
// K2: adjoint backward over the stored trajectory (replaces tf.gradients for train.py:89)
//   per-clip outputs: G[b][0]=sum_k s_k mu_k x_k^dag, G[b][1]=sum_k mu_k x_k^dag,
//                     G[b][2]=sum_k alpha_k x'_k x'_k^dag,  gf[b], lam0[b], gAdir[b]
// Adjoint recursion (DESIGN.md "Adjoint"), k descending, lam = adjoint of x_{k+1}:
//     mu_k  = c_k conj(q_k) lam + alpha_k S x'_k          alpha_k = 2 gE_k / |x_k|^2
//     lam   = L_k^dag mu_k + beta_k x_k                   beta_k  = -alpha_k E_k
// Slot c (c = nchunks .. -1):  CHAIN warps run the recursion over chunk c;  FILLER warps issue
// the loads of chunk c-2, accumulate the rank-1 gradient tiles of chunk c+1 (its mu_k are
// complete) and prepare chunk c-1 (x', S x', alpha, beta, packed per-row chain inputs).
// -------------------------------------------------------------------------------------------
template <int DP, int NQ>
__global__ void __launch_bounds__(2 * DP * NQ, 1)
    psi_bwd_kernel(const float2* __restrict__ matN, const float2* __restrict__ matRH,
                   const float2* __restrict__ matS, const float2* __restrict__ qtab,
                   const float* __restrict__ ttab, const float* __restrict__ x, int T, AVal A_,
                   const float* __restrict__ w, const float2* __restrict__ traj,
                   const float* __restrict__ scales, int nchunks, float2* __restrict__ Gout,
                   float* __restrict__ gfout, float2* __restrict__ lam0out,
                   double* __restrict__ gAdir, const float2* __restrict__ sptraj,
                   const float2* __restrict__ evin, SegBwd seg) {
  const float A = a_get(A_);
  using M = Map<DP, NQ>;
  using Sm = BwdSmem<DP, NQ>;
  constexpr int NTC = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int NP = M::NP;
  constexpr int CHK = Sm::CHK;
  constexpr int LV = (NQ == 4) ? 2 : 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm& sm = *reinterpret_cast<Sm*>(smem_raw);

  const int t = threadIdx.x;
  const bool is_chain = t >= NTC;
  const int tr = is_chain ? t - NTC : t;
  const int i = tr / NQ, jq = tr % NQ, lane = t & 31;
  const int b = blockIdx.x;
  const int nsteps = T - 1;
  auto chunk_len = [&](int c) { return min(CHK, nsteps - c * CHK); };

  if (is_chain) {
    // =================================== CHAIN WARPS ==========================================
    float2 Nr[CPT], Hr[CPT];
    load_slice<DP, NQ>(Nr, matN, i, jq);   // N is Hermitian: N^dag mu uses the same slices
    load_slice<DP, NQ>(Hr, matRH, i, jq);  // R^dag
    float2 lam = make_float2(0.f, 0.f);    // adjoint of x_{k+1}, replicated over the NQ lanes
    if (seg.lam_end) lam = seg.lam_end[(size_t)b * DP + i];
    float gf = seg.accumulate ? gfout[(size_t)b * DP + i] : 0.f;
    const bool mu_on = jq == 0;
    for (int c = nchunks; c >= -1; --c) {
      if (c >= 0 && c < nchunks) {
        const int ca = c & 1, len = chunk_len(c);
        // shared-window addresses of this chunk's chain inputs, computed once (see lds64a)
        const unsigned mu_a = smem_addr_pinned(&sm.mus[ca][0][i]);
        const unsigned cina_a = smem_addr_pinned(&sm.cina[ca][0][i]);
        const unsigned cinb_a = smem_addr_pinned(&sm.cinb[ca][0][i]);
        const unsigned sv_a = smem_addr_pinned(&sm.sv[c % 3][0]);
        constexpr unsigned ROW4 = DP * sizeof(float4), ROW2 = DP * sizeof(float2);
        {  // adjoint of x' for the chunk's last step (its q carries the rescale c_k)
          const float4 a4 = lds128v(cina_a + (unsigned)(len - 1) * ROW4);
          float2 mu = cmul_ca(make_float2(a4.x, a4.y), lam);
          mu.x += a4.z;
          mu.y += a4.w;
          sts64a_if(mu_on, mu_a + (unsigned)(len - 1) * ROW2, mu);
        }
        bar_named(1, NTC);
        const unsigned mus_addr = smem_addr_pinned(&sm.mus[ca][0][2 * jq]);
        auto step = [&](int kk) {
          float2 mv[CPT];
#pragma unroll
          for (int m = 0; m < NP; ++m) {
            const float4 v = lds128v(mus_addr + (unsigned)((kk * DP + 2 * NQ * m) * sizeof(float2)));
            mv[2 * m] = make_float2(v.x, v.y);
            mv[2 * m + 1] = make_float2(v.z, v.w);
          }
          const unsigned km = (unsigned)(kk > 0 ? kk - 1 : 0);
          const float s = lds32a(sv_a + (unsigned)kk * (unsigned)sizeof(float));
          const float4 b4 = lds128v(cinb_a + (unsigned)kk * ROW4);   // { beta x_k , dtm x_k }
          const float4 a4 = lds128v(cina_a + km * ROW4);             // { c q , alpha S x' } of step kk-1
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
          gf = fmaf(lam.x, b4.w, fmaf(-lam.y, b4.z, gf));   // dtm Im(conj(lam) x_k)
          float2 mu = cmul_ca(make_float2(a4.x, a4.y), lam);
          mu.x += a4.z;
          mu.y += a4.w;
          sts64a_if(mu_on && kk > 0, mu_a + km * ROW2, mu);
          bar_named(1, NTC);
        };
        if (len == CHK) {
#pragma unroll 2
          for (int kk = CHK - 1; kk >= 0; --kk) step(kk);
        } else {
          for (int kk = len - 1; kk >= 0; --kk) step(kk);
        }
      }
      __syncthreads();   // slot hand-over
    }
    if (jq == 0) {
      gfout[(size_t)b * DP + i] = gf;
      lam0out[(size_t)b * DP + i] = lam;
    }
  } else {
    // =================================== FILLER WARPS =========================================
    const float* xb = x + (size_t)b * seg.xstride;
    const float2* trb = traj + (size_t)b * T * DP;
    const float wb = w[b];
    float2 GR[CPT], GN[CPT], GE[CPT];
    {
      const float2* Gb = Gout + (size_t)b * 3 * DP * DP;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int col = M::col(c, jq);
        GR[c] = seg.accumulate ? Gb[0 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
        GN[c] = seg.accumulate ? Gb[1 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
        GE[c] = seg.accumulate ? Gb[2 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
      }
    }
    double gAacc = 0.0;
    const bool tpv = seg.tprev_valid != 0;

    auto issue_loads = [&](int c) {
      const int k0 = c * CHK, len = chunk_len(c);
      const float2* xsrc = trb + (size_t)k0 * DP;
      float2* xdst = &sm.xs[c & 3][0][0];
      for (int idx = tr; idx < (len + 1) * DP / 2; idx += NTC) cp_async16(xdst + 2 * idx, xsrc + 2 * idx);
      const float2* qsrc = qtab + (size_t)k0 * DP;
      float2* qdst = &sm.qs[c & 1][0][0];
      for (int idx = tr; idx < len * DP / 2; idx += NTC) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
      for (int idx = tr; idx <= len; idx += NTC) {
        cp_async4(&sm.wav[c & 1][idx], xb + k0 + idx);
        // tt[0] = t_{k0-1} (t_0 for the clip's first chunk: dtm_0 is forced to 0 there), tt[1+kk] = t_{k0+kk}
        cp_async4(&sm.tt[c & 1][idx], ttab + ((k0 + idx > 0 || tpv) ? k0 + idx - 1 : 0));
      }
      if (tr == 0) cp_async4(&sm.scs[c & 1][0], scales + (size_t)b * nchunks + c);
      const float2* ssrc = sptraj + ((size_t)b * T + k0) * DP;
      float2* sdst = &sm.spl[c & 1][0][0];
      for (int idx = tr; idx < len * DP / 2; idx += NTC) cp_async16(sdst + 2 * idx, ssrc + 2 * idx);
      const float2* esrc = evin + (size_t)b * T + k0;
      for (int idx = tr; idx < 2 * len; idx += NTC)
        cp_async4(reinterpret_cast<float*>(&sm.evl[c & 1][0]) + idx, reinterpret_cast<const float*>(esrc) + idx);
    };

    // rank-1 gradient tiles of one finished chunk (its mu_k are all in shared memory)
    auto tiles_chunk = [&](int c) {
      const int len = chunk_len(c);
      const float2(*xsb)[DP] = sm.xs[c & 3];
      const float2(*xpb)[DP] = sm.xps[c % 3];
      const float2(*mub)[DP] = sm.mus[c & 1];
      for (int kk = 0; kk < len; ++kk) {
        const float2 mui = mub[kk][i];
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

    // everything the chain needs for one chunk, from the landed trajectory (elementwise: S x'_k,
    // E_k and |x_k|^2 were stored by the forward)
    auto prep_chunk = [&](int c) {
      const int len = chunk_len(c), k0 = c * CHK;
      const int lx = c & 3, lq = c & 1, lp3 = c % 3;
      const float sc = sm.scs[lq][0];
      const float inv_sc = 1.0f / sc;
      if (tr < len) {
        const float inc = sm.wav[lq][tr + 1] - sm.wav[lq][tr];
        const float s = inc / A;
        sm.sv[lp3][tr] = s;
        sm.dtm[tr] = (k0 + tr > 0 || tpv) ? sm.tt[lq][tr + 1] - sm.tt[lq][tr] : 0.f;
        const float2 ev = sm.evl[lq][tr];
        const float E = ev.x, nu2 = ev.y;
        const float arg = 1.0f + (E * inc) / A;
        const float gE = wb * (-s / arg);
        const float alpha = 2.0f * gE / fmaxf(nu2, 1e-12f);
        sm.alphas[lp3][tr] = alpha;
        sm.betas[tr] = -alpha * E;
        gAacc += (double)wb * (double)E * (double)inc / ((double)A * (double)A * (double)arg);
      }
      bar_named(2, NTC);
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
        sm.cina[lq][kk][r] = make_float4(q.x, q.y, al * sp.x, al * sp.y);
        sm.cinb[lq][kk][r] = make_float4(be * xk.x, be * xk.y, dt * xk.x, dt * xk.y);
      }
    };

    if (nchunks > 0) issue_loads(nchunks - 1);
    cp_async_commit();
    for (int c = nchunks; c >= -1; --c) {
      if (c - 2 >= 0) issue_loads(c - 2);
      cp_async_commit();
      if (c + 1 >= 0 && c + 1 < nchunks) tiles_chunk(c + 1);
      cp_async_wait<1>();    // the loads of chunk c-1 (issued one slot earlier) have landed
      bar_named(2, NTC);
      if (c - 1 >= 0) prep_chunk(c - 1);
      __syncthreads();       // slot hand-over
    }
    cp_async_wait<0>();

    // ---- per-clip outputs -------------------------------------------------------------------
    float2* Gb = Gout + (size_t)b * 3 * DP * DP;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = M::col(c, jq);
      Gb[0 * DP * DP + i * DP + col] = GR[c];
      Gb[1 * DP * DP + i * DP + col] = GN[c];
      Gb[2 * DP * DP + i * DP + col] = GE[c];
    }
    gAacc = warp_sum_d(gAacc);
    if (lane == 0) sm.lred[tr >> 5] = gAacc;
    bar_named(2, NTC);
    if (tr == 0) {
      double tot = seg.accumulate ? gAdir[b] : 0.0;
      for (int wv = 0; wv < NTC / 32; ++wv) tot += sm.lred[wv];
      gAdir[b] = tot;
    }
  }
}

// ===========================================================================================
// Unified (non-specialised) variants
// ===========================================================================================
using TrueT = std::true_type;
using FalseT = std::false_type;
using IC0 = std::integral_constant<int, 0>;
using IC1 = std::integral_constant<int, 1>;
using IC2 = std::integral_constant<int, 2>;
using IC3 = std::integral_constant<int, 3>;

// SXO (chain-only forward): the S x' / E_k staging arrays shrink to nothing -- 85 KB instead of 118 KB, so TWO
// of these 64-register CTAs share an SM (a batch of 256 clips runs as one wave on 148 SMs instead of two)
template <int DP, int NQ, bool SXO = false>
struct alignas(16) FwdSmemUni {
  static constexpr int NT = DP * NQ, G = NT / 32;
  float2 xs[CH + 1][DP];        // x_{k0+kk}
  float2 xps[CH][DP];           // x'_{k0+kk}
  float2 qs[2][CH][DP];         // q_k, double buffered
  float2 sps[SXO ? 1 : CH][DP];           // S x'_{k0+kk} (chunk-end pass; kept for the backward)
  float esr[SXO ? 1 : CH][2 * DP + 2];    // Re x'_i Re(S x')_i and Im x'_i Im(S x')_i (even / odd lane of the row)
  float ns[2][CH + 1][DP + 1];  // |x_{k,i}|^2, by chunk parity
  float2 evs[CH];               // (E_k, |x_k|^2)
  float wav[2][CH + 4];         // waveform samples k0..k0+len, double buffered
  float sv[2][CH + 4];          // s_k
  float incv[2][CH];            // inc_k
  double lred[32];
};

// CHAIN (the chain-only sweep of the tensor-core path, TILES = false and not VIRT): inputs double- instead of triple-
// buffered, no x' reconstruction, and a two-row mu ring (mu_k goes to global memory as it is formed) -- 103 KB
// instead of 195 KB, so that two CTAs share an SM and C4's 256 clips run as ONE wave.
template <int DP, bool CHAIN = false>
struct alignas(16) BwdSmemUni {
  static constexpr int NB = CHAIN ? 2 : 3;
  static constexpr int NMU = CHAIN ? 2 : CH;
  float2 xs[NB][CH + 1][DP];  // trajectory chunk, NB-buffered (3: chunk c, c-1 in use, c-2 landing; 2: c in use, c-1 landing)
  float2 qs[NB][CH][DP];
  float2 spl[NB][CH][DP];     // S x'_k stored by the forward
  float2 evl[NB][CH];         // (E_k, |x_k|^2) stored by the forward
  float2 xps[CHAIN ? 1 : 2][CHAIN ? 1 : CH][DP];   // reconstructed x'_k (tile fillers only)
  float2 mus[NMU][DP];        // adjoint of x'_k
  float wav[NB][CH + 4];
  float tt[NB][CH + 4];
  float scs[NB][4];
  float sv[2][CH], incv[2][CH], dtk[2][CH], alphas[2][CH], betas[2][CH];
  double lred[32];
};

// -------------------------------------------------------------------------------------------
// K1u: forward per-clip loss, UNIFIED variant (every warp does everything; used where the CTA
// already has several warps per sub-partition, DP = 64)
// K1u: forward per-clip loss (model.py:257-267, 276-282, 293-334)
// -------------------------------------------------------------------------------------------
// VIRT: "virtual clip" mode of the parallel-in-time scan (amps_scan_tc.cuh): block b replays time
// chunk (b % nvc) of clip (b / nvc) -- m_steps steps from global step (b % nvc) * m_steps -- starting
// from its own state psi0v[b]; the per-chunk loss goes to lossd[b].
// SXO ("S x' offloaded"): chain only -- x'_k goes where S x'_k would (sptraj rows) and |x_k|^2 into ev[k].y;
// S x'_k, E_k and the loss are produced afterwards, in place, by psi_sx_tc_kernel (amps_sx_tc.cuh) on the
// tensor cores.  Needs the trajectory buffers (a saving forward).
#ifndef AMPS_UNI_FWD_MINB
#define AMPS_UNI_FWD_MINB 0
#endif
template <int DP, int NQ, bool VIRT = false, bool SXO = false>
#if AMPS_UNI_FWD_MINB
__global__ void __launch_bounds__(DP* NQ, AMPS_UNI_FWD_MINB)
#else
__global__ void __launch_bounds__(DP* NQ)
#endif
    psi_fwd_uni_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                   const float2* __restrict__ matS, const float2* __restrict__ qtab_,
                   const float2* __restrict__ psi0p_, const float* __restrict__ x, int T, AVal A_,
                   float* __restrict__ loss, double* __restrict__ lossd,
                   float2* __restrict__ traj, float* __restrict__ scales, int nchunks,
                   const float2* __restrict__ psi0v, int nvc, int m_steps,
                   float2* __restrict__ sptraj, float2* __restrict__ evout, SegFwd seg) {
  const float A = a_get(A_);
  using M = Map<DP, NQ>;
  using Sm = FwdSmemUni<DP, NQ, SXO>;
  constexpr int NT = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int G = Sm::G, PER = DP / G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm& sm = *reinterpret_cast<Sm*>(smem_raw);

  const int t = threadIdx.x, i = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  int nsteps = T - 1;
  const float* xb = x + (size_t)b * (VIRT ? T : seg.xstride);
  const float2* qtab = qtab_;
  const float2* psi0p = (!VIRT && seg.x0) ? seg.x0 + (size_t)b * seg.x0_stride : psi0p_;
  size_t tstride = T;        // trajectory rows / rescale factors per (virtual) clip
  int sstride = nchunks;
  if (VIRT) {
    const int clip = b / nvc, kbeg = (b % nvc) * m_steps;
    nsteps = max(0, min(m_steps, T - 1 - kbeg));
    xb = x + (size_t)clip * T + kbeg;
    qtab = qtab_ + (size_t)kbeg * DP;
    psi0p = psi0v + (size_t)b * DP;
    nchunks = (nsteps + CH - 1) / CH;
    tstride = m_steps + 1;
    sstride = m_steps / CH;
  }

  float2 Nr[CPT], Rr[CPT], Sr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);
  load_slice<DP, NQ>(Rr, matR, i, jq);
  load_slice<DP, NQ>(Sr, matS, i, jq);

  if (t < DP) {
    const float2 p = psi0p[t];
    sm.xs[0][t] = p;
    sm.ns[0][0][t] = cabs2(p);
    if (traj) traj[(size_t)b * tstride * DP + t] = p;
  }

  auto issue_loads = [&](int c, int buf) {
    const int k0 = c * CH;
    const int len = min(CH, nsteps - k0);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[buf][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NT) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = t; idx <= len; idx += NT) cp_async4(&sm.wav[buf][idx], xb + k0 + idx);
  };
  auto compute_s = [&](int buf, int len) {
    if (t < len) {
      const float inc = sm.wav[buf][t + 1] - sm.wav[buf][t];   // model.py:263
      sm.incv[buf][t] = inc;
      sm.sv[buf][t] = inc / A;                                  // model.py:303
    }
  };

  // branch-free per-step stores: lane jq==0 writes x_{k+1,i}, jq==1 writes x'_{k,i} (same row
  // stride), jq==2 writes |x_{k+1,i}|^2.  The step loop carries ONLY the chain; everything that does
  // not feed the next state (S x', E_k, the loss) runs once per chunk, barrier-free, below.
  float2* const st2 = (jq == 0) ? &sm.xs[1][i] : &sm.xps[0][i];
  const bool st2_on = jq < 2;
  const bool stn_on = jq == 2;

  double lossacc = 0.0;
  if (nchunks > 0) {
    issue_loads(0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    compute_s(0, min(CH, nsteps));
  }

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    const int k0 = c * CH;
    const int len = min(CH, nsteps - k0);
    if (c + 1 < nchunks) issue_loads(c + 1, buf ^ 1);
    cp_async_commit();
    __syncthreads();  // (D) sv/incv, xs[0], ns[.][0] of this chunk visible; last chunk's flush done
    if (!VIRT) if (seg.ckpt && c % seg.ck_chunks == 0 && t < DP)   // state checkpoint: x at the start of chunk c
      seg.ckpt[(size_t)b * seg.ck_stride + (size_t)(c / seg.ck_chunks) * DP + t] = sm.xs[0][t];

    float* const stn = &sm.ns[buf][1][i];
    float2* const sp_st = &sm.sps[0][i];
    float s_cur = sm.sv[buf][0];
    float2 part_pp = make_float2(0.f, 0.f);   // per-thread partial of (S x'_{kk-2})_i, reduced in step kk

    // One step: the chain mat-vec with L_k on the critical path.  In the shadows of its loads and
    // shuffles run, software-pipelined, the expectation of EARLIER steps: (stage >= 1) the partial
    // mat-vec S x'_{kk-1}, (stage >= 2) the row reduction of the partial of step kk-2 and its stores.
    // lane 0 of a row ends with Re (S x')_i, lane 1 with Im (S x')_i (pair_reduce); each stores its
    // component of S x' and its half of Re(conj(x'_i) (S x')_i)
    float* const spf_st = reinterpret_cast<float*>(sp_st) + (jq & 1);
    float* const esf_st = &sm.esr[0][2 * i + (jq & 1)];
    const bool ex_on = jq < 2;
    auto finish_expect = [&](float2 part, int kk) {       // part: per-thread partial of (S x'_kk)_i
      const float r = pair_reduce<NQ>(part, jq);
      const float2 xpi = sm.xps[kk][i];
      sts_if(ex_on, spf_st + kk * (2 * DP), r);
      sts_if(ex_on, esf_st + kk * (2 * DP + 2), ((jq & 1) ? xpi.y : xpi.x) * r);
    };
    auto step = [&](auto stage_tag, int kk) {
      constexpr int STAGE = decltype(stage_tag)::value;
      float2 xv[CPT], pv[CPT];
#pragma unroll
      for (int m = 0; m < CPT / 2; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.xs[kk][2 * NQ * m + 2 * jq]);
        xv[2 * m] = make_float2(v.x, v.y);
        xv[2 * m + 1] = make_float2(v.z, v.w);
      }
      if (STAGE >= 1) {
#pragma unroll
        for (int m = 0; m < CPT / 2; ++m) {
          const float4 v = *reinterpret_cast<const float4*>(&sm.xps[kk - 1][2 * NQ * m + 2 * jq]);
          pv[2 * m] = make_float2(v.x, v.y);
          pv[2 * m + 1] = make_float2(v.z, v.w);
        }
      }
      const float2 q = sm.qs[buf][kk][i];
      const float s_next = sm.sv[buf][kk + 1];
      // row reduction of step kk-2's partial: three dependent shuffle levels, started first so that
      // they ride under the chain's own latency
      float red = 0.f;
      float2 xpi2 = make_float2(0.f, 0.f);
      if (STAGE >= 2) {
        xpi2 = sm.xps[kk - 2][i];
        const bool odd = jq & 1;       // pair_reduce, level 1
        red = (odd ? part_pp.y : part_pp.x) + __shfl_xor_sync(0xffffffffu, odd ? part_pp.x : part_pp.y, 1);
      }
      float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
      for (int cc = 0; cc < CPT; cc += 2) {
        const float2 l0 = make_float2(fmaf(s_cur, Rr[cc].x, Nr[cc].x), fmaf(s_cur, Rr[cc].y, Nr[cc].y));
        const float2 l1 = make_float2(fmaf(s_cur, Rr[cc + 1].x, Nr[cc + 1].x), fmaf(s_cur, Rr[cc + 1].y, Nr[cc + 1].y));
        cmac(a0, l0, xv[cc]);
        cmac(a1, l1, xv[cc + 1]);
      }
      float2 xp = make_float2(a0.x + a1.x, a0.y + a1.y);
      if (STAGE >= 2) red += __shfl_xor_sync(0xffffffffu, red, 2);
      // chain row reduction, the S x'_{kk-1} FMAs in the shuffle shadows
      float2 p0 = make_float2(0.f, 0.f), p1 = p0;
      constexpr int LV = (NQ == 4) ? 2 : 3;
      constexpr int CPL = (CPT + LV - 1) / LV;
#pragma unroll
      for (int lv = 0; lv < LV; ++lv) {
        const float ox = __shfl_xor_sync(0xffffffffu, xp.x, 1 << lv);
        const float oy = __shfl_xor_sync(0xffffffffu, xp.y, 1 << lv);
        if (STAGE >= 1) {
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
      sts_if(st2_on, st2 + kk * DP, (jq == 0) ? xn : xp);
      sts_if(stn_on, stn + kk * (DP + 1), cabs2(xn));
      if (STAGE >= 2) {
        if (NQ == 8) red += __shfl_xor_sync(0xffffffffu, red, 4);
        sts_if(ex_on, spf_st + (kk - 2) * (2 * DP), red);
        sts_if(ex_on, esf_st + (kk - 2) * (2 * DP + 2), ((jq & 1) ? xpi2.y : xpi2.x) * red);
      }
      if (STAGE >= 1) part_pp = make_float2(p0.x + p1.x, p0.y + p1.y);
      s_cur = s_next;
      __syncthreads();
    };

    if (SXO) {
      if (len == CH) {
#pragma unroll 2
        for (int kk = 0; kk < CH; ++kk) step(IC0{}, kk);
      } else {
        for (int kk = 0; kk < len; ++kk) step(IC0{}, kk);
      }
      cp_async_wait<0>();
    } else {
      step(IC0{}, 0);
      if (len > 1) step(IC1{}, 1);
      if (len == CH) {
#pragma unroll 2
        for (int kk = 2; kk < CH; ++kk) step(IC2{}, kk);
      } else {
        for (int kk = 2; kk < len; ++kk) step(IC2{}, kk);
      }
      cp_async_wait<0>();
      // drain the expectation pipeline: step len-2's partial is in part_pp, step len-1 has none yet
      if (len >= 2) finish_expect(part_pp, len - 2);
      finish_expect(matvec1<DP, NQ>(Sr, sm.xps[len - 1], jq), len - 1);
      __syncthreads();  // (A)
    }

    {  // per-step scalars, lane-parallel over the chunk: G threads per step
      const int kk = t / G, g = t % G;
      float en = 0.f, nu2 = 0.f;
      if (kk < len) {
#pragma unroll
        for (int r = 0; r < PER; ++r) {
          if (!SXO) en += sm.esr[kk][2 * (g * PER + r)] + sm.esr[kk][2 * (g * PER + r) + 1];
          nu2 += sm.ns[buf][kk][g * PER + r];
        }
      }
#pragma unroll
      for (int m = 1; m < G; m <<= 1) {
        if (!SXO) en += __shfl_xor_sync(0xffffffffu, en, m);
        nu2 += __shfl_xor_sync(0xffffffffu, nu2, m);
      }
      if (g == 0 && kk < len) {
        if (SXO) {
          sm.evs[kk] = make_float2(0.f, nu2);                        // E_k: psi_sx_tc_kernel
        } else {
          const float E = en / fmaxf(nu2, 1e-12f);                                  // model.py:324-325 on x'
          const float z = (E * sm.incv[buf][kk]) / A;                // model.py:294
          lossacc -= (double)log1pf(z);
          sm.evs[kk] = make_float2(E, nu2);
        }
      }
    }
    if (c + 1 < nchunks) compute_s(buf ^ 1, min(CH, nsteps - (k0 + CH)));
    // rescale by 1/|x_{k0+len}| (every warp computes the norm redundantly: no extra barrier)
    float n2 = 0.f;
    for (int r = lane; r < DP; r += 32) n2 += sm.ns[buf][len][r];
    n2 = warp_sum_f(n2);
    const float sc = rsqrtf(fmaxf(n2, 1e-12f));   // clamp of model.py:331-333
    if (t < DP) {
      float2 v = sm.xs[len][t];
      v.x *= sc;
      v.y *= sc;
      sm.xs[len][t] = v;
      sm.xs[0][t] = v;
      sm.ns[buf ^ 1][0][t] = cabs2(v);
    }
    if (t == 0 && scales) scales[(size_t)b * sstride + c] = sc;
    if (traj) {
      __syncthreads();  // (B) scaled x_{k0+len} visible
      const float4* src = reinterpret_cast<const float4*>(&sm.xs[1][0]);
      float4* dst = reinterpret_cast<float4*>(traj + ((size_t)b * tstride + k0 + 1) * DP);
      for (int idx = t; idx < len * DP / 2; idx += NT) dst[idx] = src[idx];
      if (sptraj) {   // S x'_k (SXO: x'_k) and (E_k, |x_k|^2) for the adjoint: row k of the (virtual) clip
        const float4* ssrc = reinterpret_cast<const float4*>(SXO ? &sm.xps[0][0] : &sm.sps[0][0]);
        float4* sdst = reinterpret_cast<float4*>(sptraj + ((size_t)b * tstride + k0) * DP);
        for (int idx = t; idx < len * DP / 2; idx += NT) sdst[idx] = ssrc[idx];
        if (t < len) evout[(size_t)b * tstride + k0 + t] = sm.evs[t];
      }
    }
  }

  // block reduction of the per-thread loss partials
  lossacc = warp_sum_d(lossacc);
  if (lane == 0) sm.lred[warp] = lossacc;
  __syncthreads();
  if (t == 0 && !SXO) {
    double tot = 0.0;
    for (int wv = 0; wv < NT / 32; ++wv) tot += sm.lred[wv];
    if (loss) loss[b] = (float)tot;
    if (lossd) lossd[b] = tot;
  }
}

// -------------------------------------------------------------------------------------------
// K2u: adjoint backward, UNIFIED variant
// K2: adjoint backward over the stored trajectory (replaces tf.gradients for train.py:89)
//   per-clip outputs: G[b][0]=sum_k s_k mu_k x_k^dag, G[b][1]=sum_k mu_k x_k^dag,
//                     G[b][2]=sum_k alpha_k x'_k x'_k^dag,  gf[b], lam0[b], gAdir[b]
// Adjoint recursion (DESIGN.md "Adjoint"), k descending, lam = adjoint of x_{k+1}:
//     mu_k  = c_k conj(q_k) lam + alpha_k S x'_k          alpha_k = 2 gE_k / |x_k|^2
//     lam   = L_k^dag mu_k + beta_k x_k                   beta_k  = -alpha_k E_k
// Per loop iteration kk of chunk c: chain mat-vec of step kk (critical path) with, as filler, the
// rank-1 tile updates of step kk and the S x' mat-vec of step kk of chunk c-1.
// -------------------------------------------------------------------------------------------
// VIRT: virtual-clip mode of the parallel-in-time scan: block b runs the adjoint of time chunk
// (b % nvc) of clip (b / nvc) over the trajectory the VIRT forward stored for it, starting from the
// adjoint lam_end[b] of the chunk's (normalised) end state (zero when lam_end is NULL).
// TILES = false: adjoint chain only (no gradient tiles) -- the first pass of the scan backward, which
// needs nothing but the adjoint of each chunk's start state.
template <int DP, int NQ, bool VIRT = false, bool TILES = true>
__global__ void __launch_bounds__(DP* NQ)
    psi_bwd_uni_kernel(const float2* __restrict__ matN, const float2* __restrict__ matRH,
                   const float2* __restrict__ matS, const float2* __restrict__ qtab_,
                   const float* __restrict__ ttab_, const float* __restrict__ x, int T, AVal A_,
                   const float* __restrict__ w, const float2* __restrict__ traj,
                   const float* __restrict__ scales_, int nchunks, float2* __restrict__ Gout,
                   float* __restrict__ gfout, float2* __restrict__ lam0out,
                   double* __restrict__ gAdir, const float2* __restrict__ lam_end, int nvc,
                   int m_steps, const float2* __restrict__ sptraj, const float2* __restrict__ evin,
                   SegBwd seg, float2* __restrict__ mu_out = nullptr) {
  // mu_out (chain-only instantiation): the adjoint mu_k of x'_k goes to row k of mu_out[b] for the
  // tensor-core tile kernel (amps_tiles_tc.cuh).  It may alias sptraj: a chunk's S x' rows have been read
  // into shared memory (cp.async, two chunks ahead) before its mu rows are written.
  const float A = a_get(A_);
  using M = Map<DP, NQ>;
  constexpr int NT = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int NP = M::NP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool CHAINM = !TILES && !VIRT;   // chain-only sweep: compact shared memory (BwdSmemUni)
  using Sm = BwdSmemUni<DP, CHAINM>;
  constexpr int NB = Sm::NB;
  Sm& sm = *reinterpret_cast<Sm*>(smem_raw);

  const int t = threadIdx.x, i = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  int nsteps = T - 1;
  const float* xb = x + (size_t)b * (VIRT ? T : seg.xstride);
  const bool accum = !VIRT && seg.accumulate;
  size_t rows = T;                       // trajectory rows per (virtual) clip
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
    scales = scales_ + (size_t)b * (m_steps / CH);
    nchunks = (nsteps + CH - 1) / CH;
    wb = w[clip];
  } else {
    wb = w[b];
  }
  const float2* trb = traj + (size_t)b * rows * DP;
  const float2* spb = sptraj + (size_t)b * rows * DP;
  const float2* evb = evin + (size_t)b * rows;

  float2 Nr[CPT], Hr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);   // N is Hermitian: N^dag mu uses the same slices
  load_slice<DP, NQ>(Hr, matRH, i, jq);  // R^dag

  float2 GR[CPT], GN[CPT], GE[CPT];
  {
    const float2* Gb = Gout + (size_t)b * 3 * DP * DP;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = M::col(c, jq);
      GR[c] = (TILES && accum) ? Gb[0 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
      GN[c] = (TILES && accum) ? Gb[1 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
      GE[c] = (TILES && accum) ? Gb[2 * DP * DP + i * DP + col] : make_float2(0.f, 0.f);
    }
  }

  auto chunk_len = [&](int c) { return min(CH, nsteps - c * CH); };

  auto issue_loads = [&](int c) {
    const int lb = c % NB;
    const int k0 = c * CH;
    const int len = chunk_len(c);
    const float2* xsrc = trb + (size_t)k0 * DP;
    float2* xdst = &sm.xs[lb][0][0];
    for (int idx = t; idx < (len + 1) * DP / 2; idx += NT) cp_async16(xdst + 2 * idx, xsrc + 2 * idx);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[lb][0][0];
    const float2* ssrc = spb + (size_t)k0 * DP;
    float2* sdst = &sm.spl[lb][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NT) {
      cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
      cp_async16(sdst + 2 * idx, ssrc + 2 * idx);
    }
    for (int idx = t; idx <= len; idx += NT) {
      cp_async4(&sm.wav[lb][idx], xb + k0 + idx);
      cp_async4(&sm.tt[lb][idx], ttab + k0 + idx);
    }
    for (int idx = t; idx < len; idx += NT) cp_async8(&sm.evl[lb][idx], evb + k0 + idx);
    if (t == 0) cp_async4(&sm.scs[lb][0], scales + c);
  };

  double gAacc = 0.0;
  // chunk c: s, inc, dt, alpha_k, beta_k, the direct dL/dA term (from the forward's (E_k, |x_k|^2));
  // x'_k = conj(q_k) x_{k+1} / c_k
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
      if (!TILES) if (mu_out)
        const_cast<float2*>(evb)[(size_t)c * CH + t] = make_float2(alpha, t == len - 1 ? 1.0f / sm.scs[lb][0] : 1.0f);
    }
    if (TILES) {   // x'_k only feeds the G_E tile
      const float inv_sc = 1.0f / sm.scs[lb][0];
      for (int idx = t; idx < len * DP; idx += NT) {
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

  float2 lam = make_float2(0.f, 0.f);  // adjoint of x_{k+1}, replicated over the NQ lanes
  if (VIRT) {
    if (lam_end) lam = lam_end[(size_t)b * DP + i];
  } else if (seg.lam_end) {
    lam = seg.lam_end[(size_t)b * DP + i];
  }
  float gf = accum ? gfout[(size_t)b * DP + i] : 0.f;

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
  }

  float2* const mu_st = &sm.mus[0][i];
  const bool mu_on = jq == 0;

  for (int c = nchunks - 1; c >= 0; --c) {
    const int lb = c % NB, ds = c & 1;
    const int len = chunk_len(c);
    if (NB == 3) {
      if (c >= 2) issue_loads(c - 2);
      cp_async_commit();
      cp_async_wait<1>();   // chunk c-1 has landed
      __syncthreads();      // (T1) prep of chunk c (previous iteration / prologue) visible
      if (c >= 1) prep_elementwise(c - 1);
    } else {
      cp_async_wait<0>();   // chunk c has landed (issued one chunk ago)
      __syncthreads();      // ... for every thread, and everybody is done with chunk c+1's buffer
      prep_elementwise(c);
      if (c >= 1) issue_loads(c - 1);   // into the buffer chunk c+1 has just left
      cp_async_commit();
      __syncthreads();      // (T1) prep of chunk c visible
    }
    const float sc = sm.scs[lb][0];
    // chain-only sweep: mu_k is kept in a two-row ring and written to global memory as it is formed
    float2* const mu_g = (CHAINM && mu_out) ? mu_out + ((size_t)b * rows + (size_t)c * CH) * DP + i : nullptr;

    // adjoint of x' for the chunk's last step (carries the rescale c_k)
    float2 mu;
    {
      const int kk = len - 1;
      const float2 xn = sm.xs[lb][kk + 1][i];
      gf = fmaf(sm.dtk[ds][kk], lam.x * xn.y - lam.y * xn.x, gf);   // Im(conj(lam) x_{k+1})
      mu = cmul_ca(sm.qs[lb][kk][i], lam);
      const float al = sm.alphas[ds][kk];
      const float2 sp = sm.spl[lb][kk][i];
      mu.x = fmaf(al, sp.x, mu.x * sc);
      mu.y = fmaf(al, sp.y, mu.y * sc);
      if (mu_on) {
        mu_st[(CHAINM ? (kk & 1) : kk) * DP] = mu;
        if (CHAINM && mu_g) mu_g[(size_t)kk * DP] = mu;
      }
    }
    __syncthreads();      // (T2)

    auto step = [&](int kk) {
      // ---- operand loads --------------------------------------------------------------
      float2 mv[CPT];
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(&sm.mus[CHAINM ? (kk & 1) : kk][2 * NQ * m + 2 * jq]);
        mv[2 * m] = make_float2(v.x, v.y);
        mv[2 * m + 1] = make_float2(v.z, v.w);
      }
      const float s = sm.sv[ds][kk];
      const float be = sm.betas[ds][kk];
      const float2 xk = sm.xs[lb][kk][i];
      const int km = kk > 0 ? kk - 1 : 0;
      const float2 q1 = sm.qs[lb][km][i];
      const float al1 = sm.alphas[ds][km];
      const float2 sp1 = sm.spl[lb][km][i];
      const float dt1 = sm.dtk[ds][km];
      // ---- chain: lam = L_k^dag mu + beta x_k -------------------------------------------
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
      // ---- filler: rank-1 tiles of step kk (mu_k is in registers), in the shuffle shadows ----
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
      lp.x += __shfl_xor_sync(0xffffffffu, lp.x, 2);
      lp.y += __shfl_xor_sync(0xffffffffu, lp.y, 2);
      if (NQ == 8) {
        lp.x += __shfl_xor_sync(0xffffffffu, lp.x, 4);
        lp.y += __shfl_xor_sync(0xffffffffu, lp.y, 4);
      }
      lam.x = fmaf(be, xk.x, lp.x);
      lam.y = fmaf(be, xk.y, lp.y);
      // ---- adjoint of x' for step kk-1 --------------------------------------------------
      if (kk > 0) {
        gf = fmaf(dt1, lam.x * xk.y - lam.y * xk.x, gf);   // Im(conj(lam) x_k), x_k = x_{(k-1)+1}
        mu = cmul_ca(q1, lam);
        mu.x = fmaf(al1, sp1.x, mu.x);
        mu.y = fmaf(al1, sp1.y, mu.y);
      }
      sts_if(mu_on && kk > 0, mu_st + (CHAINM ? (km & 1) : km) * DP, mu);
      if (CHAINM) if (mu_on && kk > 0 && mu_g) mu_g[(size_t)km * DP] = mu;
      __syncthreads();
    };

    for (int kk = len - 1; kk >= 0; --kk) step(kk);
    if (!TILES && !CHAINM) if (mu_out) {   // flush the chunk's mu ring (complete after the last step's barrier)
      const float4* src = reinterpret_cast<const float4*>(&sm.mus[0][0]);
      float4* dst = reinterpret_cast<float4*>(mu_out + ((size_t)b * rows + (size_t)c * CH) * DP);
      for (int idx = t; idx < len * DP / 2; idx += NT) dst[idx] = src[idx];
    }
  }
  cp_async_wait<0>();

  // ---- per-clip outputs -------------------------------------------------------------------
  if (TILES) {
    float2* Gb = Gout + (size_t)b * 3 * DP * DP;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = M::col(c, jq);
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
  if (t == 0) {
    double tot = accum ? gAdir[b] : 0.0;
    for (int wv = 0; wv < NT / 32; ++wv) tot += sm.lred[wv];
    gAdir[b] = tot;
  }
}

// -------------------------------------------------------------------------------------------
// K3: sampler (model.py:242-251, 284-291) from a supplied noise tensor [L][n]
// -------------------------------------------------------------------------------------------
// (the occupancy hint matters: without it ptxas keeps the kernel at 64 registers and serialises the state loads
// behind the FFMAs that free their registers -- C2 28.0 ms; with it 93 registers, 24.0 ms)
template <int DP, int NQ>
__global__ void __launch_bounds__(DP* NQ, (DP * NQ <= 256 ? 2 : 1))
    psi_sample_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                      const float2* __restrict__ qtab, const float2* __restrict__ psi0p,
                      const float* __restrict__ noise, int L, int n, float A, float dtf,
                      float* __restrict__ out) {
  using M = Map<DP, NQ>;
  constexpr int NT = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SampleSmem<DP>& sm = *reinterpret_cast<SampleSmem<DP>*>(smem_raw);

  const int t = threadIdx.x, i = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  const int nchunks = (L + CH - 1) / CH;

  float2 Nr[CPT], Rr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);
  load_slice<DP, NQ>(Rr, matR, i, jq);
  if (t < DP) sm.xs[0][t] = psi0p[t];

  auto issue_loads = [&](int c, int buf) {
    const int k0 = c * CH;
    const int len = min(CH, L - k0);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[buf][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NT) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = t; idx < len; idx += NT)
      cp_async4(&sm.nz[buf][idx], noise + (size_t)b * L + k0 + idx);   // noise: [n][L] (transposed copy)
  };

  // shared-window addresses of the step loop's arrays, computed once (see lds64a)
  const unsigned xs_a = smem_addr_pinned(&sm.xs[0][0]), wred_a = smem_addr_pinned(&sm.wred[0][0][0]);
  const unsigned qs_a = smem_addr_pinned(&sm.qs[0][0][0]), nz_a = smem_addr_pinned(&sm.nz[0][0]);
  const unsigned outs_a = smem_addr_pinned(&sm.outs[0]);
  float X = 0.f;  // cumulative sample, replicated in every thread (identical arithmetic)
  const float invA = 1.0f / A;
  if (nchunks > 0) issue_loads(0, 0);
  cp_async_commit();

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    const int k0 = c * CH;
    const int len = min(CH, L - k0);
    if (c + 1 < nchunks) issue_loads(c + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // One step reads state buffer CUR and writes CUR ^ 1.  CUR is a compile-time constant (steps are
    // issued in pairs; every chunk but the last has an even length, so a chunk always starts at 0):
    // with a run-time buffer index the compiler formed generic pointers into shared memory and
    // re-read %cluster_ctaid (S2UR/S2R, ~40 cycles each) inside every step to convert them back
    // (ncu source page, profiles/r1_chain_source.md).
    const unsigned qs_b = qs_a + (unsigned)(buf * CH * DP * sizeof(float2));
    const unsigned nz_b = nz_a + (unsigned)(buf * CH * sizeof(float));
    auto step = [&](auto cur_tag, int kk) {
      constexpr int CUR = decltype(cur_tag)::value;
      constexpr unsigned XO = CUR * DP * sizeof(float2), XN = (CUR ^ 1) * DP * sizeof(float2);
      constexpr unsigned WO = CUR * 32 * 2 * sizeof(float);
      // a = N x, y = R x (partial over this thread's columns), state read with ordered 128-bit loads
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
      for (int m = 0; m < CPT / 2; ++m) {
        const float4 xv = lds128v(xs_a + XO + (unsigned)((2 * NQ * m + 2 * jq) * sizeof(float2)));
        const float2 x0 = make_float2(xv.x, xv.y), x1 = make_float2(xv.z, xv.w);
        cmac(a0, Nr[2 * m], x0);
        cmac(a1, Nr[2 * m + 1], x1);
        cmac(b0, Rr[2 * m], x0);
        cmac(b1, Rr[2 * m + 1], x1);
      }
      float2 a = make_float2(a0.x + a1.x, a0.y + a1.y), y = make_float2(b0.x + b1.x, b0.y + b1.y);
      a = group_sum_fast<NQ>(a);
      y = group_sum_fast<NQ>(y);
      const float2 xi = lds64a(xs_a + XO + (unsigned)(i * sizeof(float2)));
      // <x, R x> and |x|^2 over rows: values are replicated over the NQ lanes of a group
      float e = fmaf(xi.x, y.x, xi.y * y.y);
      float nn = cabs2(xi);
#pragma unroll
      for (int m = NQ; m < 32; m <<= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, m);
        nn += __shfl_xor_sync(0xffffffffu, nn, m);
      }
      sts64a_if(lane == 0, wred_a + WO + (unsigned)(warp * 2 * sizeof(float)), make_float2(e, nn));
      __syncthreads();
      float es = 0.f, nsum = 0.f;
#pragma unroll
      for (int wv = 0; wv < NW; ++wv) {
        const float2 w2 = lds64a(wred_a + WO + (unsigned)(wv * 2 * sizeof(float)));
        es += w2.x;
        nsum += w2.y;
      }
      // (both divisions sit on the step's dependency chain: MUFU reciprocal + multiply, 1-2 ulp from the IEEE quotient)
      const float nsc = fmaxf(nsum, 1e-12f);
      const float E = __fdividef(2.0f * es, nsc);                                       // model.py:319-325
      const float inc = __fadd_rn(__fmul_rn(E, dtf), lds32a(nz_b + (unsigned)(kk * sizeof(float))));   // model.py:286
      X = __fadd_rn(X, inc);                                             // model.py:287
      const float s = inc * invA;                                        // model.py:303
      float rn;   // lagged normalisation keeps |x| ~ 1 (nsc >= 1e-12 is a normal number: no denormal scaling path)
      asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rn) : "f"(nsc));
      float2 xp = make_float2(fmaf(s, y.x, a.x) * rn, fmaf(s, y.y, a.y) * rn);
      const float2 xn = cmul(lds64a(qs_b + (unsigned)((kk * DP + i) * sizeof(float2))), xp);
      sts64a_if(jq == 0, xs_a + XN + (unsigned)(i * sizeof(float2)), xn);
      sts32a_if(t == 0, outs_a + (unsigned)(kk * sizeof(float)), A * X);   // model.py:251
      __syncthreads();
    };
    for (int kk = 0; kk < len; kk += 2) {
      step(IC0{}, kk);
      if (kk + 1 < len) step(IC1{}, kk + 1);
    }
    if (t < len) out[(size_t)b * L + k0 + t] = sm.outs[t];
    // outs is rewritten only after the next chunk's first two barriers
  }
  cp_async_wait<0>();
}

}  // namespace amps
