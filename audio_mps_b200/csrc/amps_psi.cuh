// PsiCMPS scan kernels (sm_100a): sequential persistent forward-loss kernel, adjoint backward,
// sampler.  One CTA owns one clip for the whole clip; the step operators live in registers,
// the state in shared memory, the waveform / phase table / trajectory stream in through
// multi-buffered cp.async.
//
// Formulation (validated against the op-for-op oracle, see DESIGN.md "Chain form"):
// in the interaction frame x_k = psi_k * conj(p_k) the reference step (model.py:276-334) is
//     x'_k    = N x_k + s_k R x_k,           N = I - (delta_t sigma^2 / 2) R^dag R,  s_k = inc_k / A
//     E_k     = x'_k^dag (R + R^dag) x'_k / |x_k|^2
//     loss   += -log(1 + (E_k inc_k) / A)
//     x_{k+1} = q_k * x'_k (* c),             q_k = p_k conj(p_{k+1})
// with p_k = exp(i fl32(f t_k)) and t_k the float32 running sum.  The state is carried
// UN-normalised (the loss is scale invariant) and rescaled by c once per chunk, so the only
// thing on the per-step critical path is one stacked [N;R] mat-vec.  Everything that does not
// feed the next state (E_k, |x_k|^2, the log, the parameter-gradient tiles) is software-pipelined
// into the same loop one step (or one chunk) behind, where it fills the issue slots the
// latency-bound chain leaves empty, or is done lane-parallel over the CH steps of a chunk.
#pragma once
#include <type_traits>

#include "amps_common.cuh"

namespace amps {

using TrueT = std::true_type;
using FalseT = std::false_type;

// -------------------------------------------------------------------------------------------
// shared-memory layouts
// -------------------------------------------------------------------------------------------
template <int DP>
struct alignas(16) FwdSmem {
  float2 xs[CH + 1][DP];     // x_{k0+kk}
  float2 xps[CH][DP];        // x'_{k0+kk}
  float2 qs[2][CH][DP];      // q_k, double buffered
  float es[CH][DP + 1];      // Re(conj(x'_i) (S x')_i)
  float ns[2][CH + 1][DP + 1];  // |x_{k,i}|^2, chunk parity (row 0 of the next chunk is written early)
  float wav[2][CH + 4];      // waveform samples k0..k0+len, double buffered
  float sv[2][CH];           // s_k
  float incv[2][CH];         // inc_k
  double lred[32];
};

template <int DP>
struct alignas(16) BwdSmem {
  float2 xs[3][CH + 1][DP];  // trajectory chunk, triple buffered (chunk c, c-1 in use, c-2 landing)
  float2 qs[3][CH][DP];
  float2 xps[2][CH][DP];     // reconstructed x'_k      (chunk c and, being prepared, c-1)
  float2 sps[2][CH][DP];     // S x'_k
  float2 mus[CH][DP];        // adjoint of x'_k
  float es[CH][DP + 1];
  float ns[CH][DP + 1];
  float wav[3][CH + 4];
  float tt[3][CH + 4];
  float scs[3][4];
  float sv[2][CH], incv[2][CH], dtk[2][CH], alphas[2][CH], betas[2][CH];
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

// Per-step scalars, lane-parallel over the chunk: G = NT/32 threads share one step, each sums
// DP/G entries of es / ns, finished with xor-shuffles.  Returns (sum es, sum ns) on every thread
// of the group; kk is the step this thread works on.
template <int DP, int NT>
__device__ __forceinline__ void chunk_scalars(const float (*es)[DP + 1], const float (*ns)[DP + 1],
                                              int len, int t, int& kk, bool& leader, float& en,
                                              float& nu2) {
  constexpr int G = NT / 32, PER = DP / G;
  kk = t / G;
  const int g = t % G;
  leader = (g == 0) && (kk < len);
  en = 0.f;
  nu2 = 0.f;
  if (kk < len) {
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      en += es[kk][g * PER + r];
      nu2 += ns[kk][g * PER + r];
    }
  }
#pragma unroll
  for (int m = 1; m < G; m <<= 1) {
    en += __shfl_xor_sync(0xffffffffu, en, m);
    nu2 += __shfl_xor_sync(0xffffffffu, nu2, m);
  }
}

// -------------------------------------------------------------------------------------------
// K1: forward per-clip loss (model.py:257-267, 276-282, 293-334)
// -------------------------------------------------------------------------------------------
template <int DP, int NQ>
__global__ void __launch_bounds__(DP* NQ)
    psi_fwd_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                   const float2* __restrict__ matS, const float2* __restrict__ qtab,
                   const float2* __restrict__ psi0p, const float* __restrict__ x, int T, float A,
                   float* __restrict__ loss, double* __restrict__ lossd,
                   float2* __restrict__ traj, float* __restrict__ scales, int nchunks) {
  using M = Map<DP, NQ>;
  constexpr int NT = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int NP = M::NP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem<DP>& sm = *reinterpret_cast<FwdSmem<DP>*>(smem_raw);

  const int t = threadIdx.x, i = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  const int nsteps = T - 1;
  const float* xb = x + (size_t)b * T;

  float2 Nr[CPT], Rr[CPT], Sr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);
  load_slice<DP, NQ>(Rr, matR, i, jq);
  load_slice<DP, NQ>(Sr, matS, i, jq);

  if (t < DP) {
    const float2 p = psi0p[t];
    sm.xs[0][t] = p;
    sm.ns[0][0][t] = cabs2(p);
    if (traj) traj[(size_t)b * T * DP + t] = p;
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
    __syncthreads();  // (D) sv/incv, xs[0], ns[0] of this chunk visible; last chunk's flush done

    float2 xp_prev = make_float2(0.f, 0.f);
    // One step of the chain (critical path) with the expectation mat-vec of the PREVIOUS step
    // software-pipelined into it.
    auto step = [&](auto first_tag, int kk) {
      constexpr bool FIRST = decltype(first_tag)::value;
      const float s = sm.sv[buf][kk];
      const float2 q = sm.qs[buf][kk][i];
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, y0 = a0, y1 = a0, p0 = a0, p1 = a0;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const float4 xv = *reinterpret_cast<const float4*>(&sm.xs[kk][2 * NQ * m + 2 * jq]);
        const float2 x0 = make_float2(xv.x, xv.y), x1 = make_float2(xv.z, xv.w);
        cmac(a0, Nr[2 * m], x0);
        cmac(a1, Nr[2 * m + 1], x1);
        cmac(y0, Rr[2 * m], x0);
        cmac(y1, Rr[2 * m + 1], x1);
        if (!FIRST) {
          const float4 pv = *reinterpret_cast<const float4*>(&sm.xps[kk - 1][2 * NQ * m + 2 * jq]);
          cmac(p0, Sr[2 * m], make_float2(pv.x, pv.y));
          cmac(p1, Sr[2 * m + 1], make_float2(pv.z, pv.w));
        }
      }
      float2 xp = make_float2(fmaf(s, y0.x + y1.x, a0.x + a1.x), fmaf(s, y0.y + y1.y, a0.y + a1.y));
      float e = 0.f;
      if (!FIRST) e = fmaf(xp_prev.x, p0.x + p1.x, xp_prev.y * (p0.y + p1.y));
#pragma unroll
      for (int m = 1; m < NQ; m <<= 1) {
        xp.x += __shfl_xor_sync(0xffffffffu, xp.x, m);
        xp.y += __shfl_xor_sync(0xffffffffu, xp.y, m);
        if (!FIRST) e += __shfl_xor_sync(0xffffffffu, e, m);
      }
      const float2 xn = cmul(q, xp);
      if (jq == 0) sm.xs[kk + 1][i] = xn;
      if (jq == 1) sm.xps[kk][i] = xp;
      if (jq == 2) sm.ns[buf][kk + 1][i] = cabs2(xn);
      if (!FIRST && jq == 3) sm.es[kk - 1][i] = e;
      xp_prev = xp;
      __syncthreads();
    };

    step(TrueT{}, 0);
    if (len == CH) {
#pragma unroll 2
      for (int kk = 1; kk < CH; ++kk) step(FalseT{}, kk);
    } else {
      for (int kk = 1; kk < len; ++kk) step(FalseT{}, kk);
    }
    {  // expectation of the chunk's last step
      const float2 part = matvec1<DP, NQ>(Sr, sm.xps[len - 1], jq);
      float e = fmaf(xp_prev.x, part.x, xp_prev.y * part.y);
#pragma unroll
      for (int m = 1; m < NQ; m <<= 1) e += __shfl_xor_sync(0xffffffffu, e, m);
      if (jq == 3) sm.es[len - 1][i] = e;
    }
    cp_async_wait<0>();
    __syncthreads();  // (A)

    {  // per-step scalars, lane-parallel over the chunk
      int kk;
      bool leader;
      float en, nu2;
      chunk_scalars<DP, NT>(sm.es, sm.ns[buf], len, t, kk, leader, en, nu2);
      if (leader) {
        const float E = en / nu2;                                  // model.py:324-325 on x'
        const float z = (E * sm.incv[buf][kk]) / A;                // model.py:294
        lossacc -= (double)log1pf(z);
      }
    }
    if (c + 1 < nchunks) compute_s(buf ^ 1, min(CH, nsteps - (k0 + CH)));
    // rescale by 1/|x_{k0+len}| (every warp computes the norm redundantly: no extra barrier)
    float n2 = 0.f;
    for (int r = lane; r < DP; r += 32) n2 += sm.ns[buf][len][r];
    n2 = warp_sum_f(n2);
    const float sc = rsqrtf(n2);
    if (t < DP) {
      float2 v = sm.xs[len][t];
      v.x *= sc;
      v.y *= sc;
      sm.xs[len][t] = v;
      sm.xs[0][t] = v;
      sm.ns[buf ^ 1][0][t] = cabs2(v);
    }
    if (t == 0 && scales) scales[(size_t)b * nchunks + c] = sc;
    if (traj) {
      __syncthreads();  // (B) scaled x_{k0+len} visible
      const float4* src = reinterpret_cast<const float4*>(&sm.xs[1][0]);
      float4* dst = reinterpret_cast<float4*>(traj + ((size_t)b * T + k0 + 1) * DP);
      for (int idx = t; idx < len * DP / 2; idx += NT) dst[idx] = src[idx];
    }
  }

  // block reduction of the per-thread loss partials
  lossacc = warp_sum_d(lossacc);
  if (lane == 0) sm.lred[warp] = lossacc;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int wv = 0; wv < NT / 32; ++wv) tot += sm.lred[wv];
    loss[b] = (float)tot;
    if (lossd) lossd[b] = tot;
  }
}

// -------------------------------------------------------------------------------------------
// K2: adjoint backward over the stored trajectory (replaces tf.gradients for train.py:89)
//   per-clip outputs: G[b][0]=sum_k s_k mu_k x_k^dag, G[b][1]=sum_k mu_k x_k^dag,
//                     G[b][2]=sum_k alpha_k x'_k x'_k^dag,  gf[b], lam0[b], gAdir[b]
// Adjoint recursion (DESIGN.md "Adjoint"), k descending, lam = adjoint of x_{k+1}:
//     mu_k  = c_k conj(q_k) lam + alpha_k S x'_k          alpha_k = 2 gE_k / |x_k|^2
//     lam   = N mu_k + s_k R^dag mu_k + beta_k x_k        beta_k  = -alpha_k E_k
// Pipeline per loop iteration kk of chunk c: chain step kk (critical path), the rank-1 tile
// updates of step kk+1, and the S x' mat-vec of step kk of chunk c-1.
// -------------------------------------------------------------------------------------------
template <int DP, int NQ>
__global__ void __launch_bounds__(DP* NQ)
    psi_bwd_kernel(const float2* __restrict__ matN, const float2* __restrict__ matRH,
                   const float2* __restrict__ matS, const float2* __restrict__ qtab,
                   const float* __restrict__ ttab, const float* __restrict__ x, int T, float A,
                   const float* __restrict__ w, const float2* __restrict__ traj,
                   const float* __restrict__ scales, int nchunks, float2* __restrict__ Gout,
                   float* __restrict__ gfout, float2* __restrict__ lam0out,
                   double* __restrict__ gAdir) {
  using M = Map<DP, NQ>;
  constexpr int NT = M::NT;
  constexpr int CPT = M::CPT;
  constexpr int NP = M::NP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem<DP>& sm = *reinterpret_cast<BwdSmem<DP>*>(smem_raw);

  const int t = threadIdx.x, i = t / NQ, jq = t % NQ, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.x;
  const int nsteps = T - 1;
  const float* xb = x + (size_t)b * T;
  const float2* trb = traj + (size_t)b * T * DP;
  const float wb = w[b];

  float2 Nr[CPT], Hr[CPT], Sr[CPT];
  load_slice<DP, NQ>(Nr, matN, i, jq);   // N is Hermitian: N^dag mu uses the same slices
  load_slice<DP, NQ>(Hr, matRH, i, jq);  // R^dag
  load_slice<DP, NQ>(Sr, matS, i, jq);

  float2 GR[CPT], GN[CPT], GE[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) GR[c] = GN[c] = GE[c] = make_float2(0.f, 0.f);

  auto chunk_len = [&](int c) { return min(CH, nsteps - c * CH); };

  auto issue_loads = [&](int c) {
    const int lb = c % 3;
    const int k0 = c * CH;
    const int len = chunk_len(c);
    const float2* xsrc = trb + (size_t)k0 * DP;
    float2* xdst = &sm.xs[lb][0][0];
    for (int idx = t; idx < (len + 1) * DP / 2; idx += NT) cp_async16(xdst + 2 * idx, xsrc + 2 * idx);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[lb][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NT) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = t; idx <= len; idx += NT) {
      cp_async4(&sm.wav[lb][idx], xb + k0 + idx);
      cp_async4(&sm.tt[lb][idx], ttab + k0 + idx);
    }
    if (t == 0) cp_async4(&sm.scs[lb][0], scales + (size_t)b * nchunks + c);
  };

  // P0 + P1 of chunk c: s, inc, dt; x'_k = conj(q_k) x_{k+1} / c_k ; |x_k|^2
  auto prep_elementwise = [&](int c) {
    const int lb = c % 3, ds = c & 1, len = chunk_len(c);
    if (t < len) {
      const float inc = sm.wav[lb][t + 1] - sm.wav[lb][t];
      sm.incv[ds][t] = inc;
      sm.sv[ds][t] = inc / A;
      sm.dtk[ds][t] = sm.tt[lb][t + 1] - sm.tt[lb][t];
    }
    const float inv_sc = 1.0f / sm.scs[lb][0];
    for (int idx = t; idx < len * DP; idx += NT) {
      const int kk = idx / DP, r = idx % DP;
      float2 xp = cmul_ca(sm.qs[lb][kk][r], sm.xs[lb][kk + 1][r]);
      if (kk == len - 1) {
        xp.x *= inv_sc;
        xp.y *= inv_sc;
      }
      sm.xps[ds][kk][r] = xp;
      sm.ns[kk][r] = cabs2(sm.xs[lb][kk][r]);
    }
  };
  // P2 of one step: S x' (stored) and e_i
  auto expectation_step = [&](int ds, int kk) {
    float2 part = matvec1<DP, NQ>(Sr, sm.xps[ds][kk], jq);
    part = group_sum<NQ>(part);
    if (jq == 0) sm.sps[ds][kk][i] = part;
    if (jq == 1) {
      const float2 xpi = sm.xps[ds][kk][i];
      sm.es[kk][i] = fmaf(xpi.x, part.x, xpi.y * part.y);
    }
  };
  double gAacc = 0.0;
  // P3 of chunk c: alpha_k, beta_k and the direct dL/dA term
  auto prep_scalars = [&](int c) {
    const int ds = c & 1, len = chunk_len(c);
    int kk;
    bool leader;
    float en, nu2;
    chunk_scalars<DP, NT>(sm.es, sm.ns, len, t, kk, leader, en, nu2);
    if (leader) {
      const float E = en / nu2;
      const float inc = sm.incv[ds][kk];
      const float arg = 1.0f + (E * inc) / A;
      const float gE = wb * (-sm.sv[ds][kk] / arg);
      const float alpha = 2.0f * gE / nu2;
      sm.alphas[ds][kk] = alpha;
      sm.betas[ds][kk] = -alpha * E;
      gAacc += (double)wb * (double)E * (double)inc / ((double)A * (double)A * (double)arg);
    }
  };

  float2 lam = make_float2(0.f, 0.f);  // adjoint of x_{k+1}, replicated over the NQ lanes
  float gf = 0.f;

  if (nchunks > 0) {
    const int cl = nchunks - 1;
    issue_loads(cl);
    cp_async_commit();
    if (cl >= 1) issue_loads(cl - 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    prep_elementwise(cl);
    __syncthreads();
    for (int kk = 0; kk < chunk_len(cl); ++kk) expectation_step(cl & 1, kk);
    __syncthreads();
    prep_scalars(cl);
  }

  for (int c = nchunks - 1; c >= 0; --c) {
    const int lb = c % 3, ds = c & 1;
    const int len = chunk_len(c);
    const bool has_prev = c >= 1;
    const int pds = ds ^ 1;
    if (c >= 2) issue_loads(c - 2);
    cp_async_commit();
    cp_async_wait<1>();   // chunk c-1 has landed
    __syncthreads();      // (T1)
    if (has_prev) prep_elementwise(c - 1);
    __syncthreads();      // (T2)
    const float sc = sm.scs[lb][0];

    float2 mu_prev = make_float2(0.f, 0.f);
    // tile update of step kk (its mu is mu_i, row values replicated in the group)
    auto tiles = [&](int kk, float2 mui) {
      const float2 xpi = sm.xps[ds][kk][i];
      const float s = sm.sv[ds][kk];
      const float al = sm.alphas[ds][kk];
      const float2 u1 = make_float2(s * mui.x, s * mui.y);
      const float2 u3 = make_float2(al * xpi.x, al * xpi.y);
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const float4 xv = *reinterpret_cast<const float4*>(&sm.xs[lb][kk][2 * NQ * m + 2 * jq]);
        const float4 pv = *reinterpret_cast<const float4*>(&sm.xps[ds][kk][2 * NQ * m + 2 * jq]);
        const float2 x0 = make_float2(xv.x, xv.y), x1 = make_float2(xv.z, xv.w);
        const float2 p0 = make_float2(pv.x, pv.y), p1 = make_float2(pv.z, pv.w);
        cmac_cx(GR[2 * m], u1, x0);
        cmac_cx(GR[2 * m + 1], u1, x1);
        cmac_cx(GN[2 * m], mui, x0);
        cmac_cx(GN[2 * m + 1], mui, x1);
        cmac_cx(GE[2 * m], u3, p0);
        cmac_cx(GE[2 * m + 1], u3, p1);
      }
    };

    auto step = [&](auto first_tag, auto prev_tag, int kk) {
      constexpr bool FIRST = decltype(first_tag)::value;   // first iteration of the chunk
      constexpr bool PREV = decltype(prev_tag)::value;     // chunk c-1 exists and has step kk
      const float2 q = sm.qs[lb][kk][i];
      const float2 xn = sm.xs[lb][kk + 1][i];
      gf = fmaf(sm.dtk[ds][kk], lam.x * xn.y - lam.y * xn.x, gf);   // Im(conj(lam) x_{k+1})
      float2 mu = cmul_ca(q, lam);
      if (FIRST) {
        mu.x *= sc;
        mu.y *= sc;
      }
      const float al = sm.alphas[ds][kk];
      const float2 sp = sm.sps[ds][kk][i];
      mu.x = fmaf(al, sp.x, mu.x);
      mu.y = fmaf(al, sp.y, mu.y);
      if (jq == 0) sm.mus[kk][i] = mu;
      __syncthreads();
      float2 a, h;
      matvec2<DP, NQ>(Nr, Hr, sm.mus[kk], jq, a, h);
      const float s = sm.sv[ds][kk];
      float2 lp = make_float2(fmaf(s, h.x, a.x), fmaf(s, h.y, a.y));
      if (!FIRST) tiles(kk + 1, mu_prev);
      if (PREV) expectation_step(pds, kk);
      lp = group_sum<NQ>(lp);
      const float be = sm.betas[ds][kk];
      const float2 xk = sm.xs[lb][kk][i];
      lam.x = fmaf(be, xk.x, lp.x);
      lam.y = fmaf(be, xk.y, lp.y);
      mu_prev = mu;
    };

    if (has_prev) {
      step(TrueT{}, TrueT{}, len - 1);
      for (int kk = len - 2; kk >= 0; --kk) step(FalseT{}, TrueT{}, kk);
      // chunk c-1 is always full; finish its expectation steps if this chunk was short
      for (int kk = len; kk < CH; ++kk) expectation_step(pds, kk);
    } else {
      step(TrueT{}, FalseT{}, len - 1);
      for (int kk = len - 2; kk >= 0; --kk) step(FalseT{}, FalseT{}, kk);
    }
    tiles(0, mu_prev);
    __syncthreads();      // (E1) es / sps of chunk c-1 complete; mus free
    if (has_prev) prep_scalars(c - 1);
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
  if (jq == 0) {
    gfout[(size_t)b * DP + i] = gf;
    lam0out[(size_t)b * DP + i] = lam;
  }
  gAacc = warp_sum_d(gAacc);
  if (lane == 0) sm.lred[warp] = gAacc;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int wv = 0; wv < NT / 32; ++wv) tot += sm.lred[wv];
    gAdir[b] = tot;
  }
}

// -------------------------------------------------------------------------------------------
// K3: sampler (model.py:242-251, 284-291) from a supplied noise tensor [L][n]
// -------------------------------------------------------------------------------------------
template <int DP, int NQ>
__global__ void __launch_bounds__(DP* NQ)
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
      cp_async4(&sm.nz[buf][idx], noise + (size_t)(k0 + idx) * n + b);
  };

  float X = 0.f;  // cumulative sample, replicated in every thread (identical arithmetic)
  int cur = 0;
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

    for (int kk = 0; kk < len; ++kk) {
      float2 a, y;
      matvec2<DP, NQ>(Nr, Rr, sm.xs[cur], jq, a, y);
      a = group_sum<NQ>(a);
      y = group_sum<NQ>(y);
      const float2 xi = sm.xs[cur][i];
      // <x, R x> and |x|^2 over rows: values are replicated over the NQ lanes of a group
      float e = fmaf(xi.x, y.x, xi.y * y.y);
      float nn = cabs2(xi);
#pragma unroll
      for (int m = NQ; m < 32; m <<= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, m);
        nn += __shfl_xor_sync(0xffffffffu, nn, m);
      }
      const int par = kk & 1;
      if (lane == 0) {
        sm.wred[par][warp][0] = e;
        sm.wred[par][warp][1] = nn;
      }
      __syncthreads();
      float es = 0.f, nsum = 0.f;
#pragma unroll
      for (int wv = 0; wv < NW; ++wv) {
        es += sm.wred[par][wv][0];
        nsum += sm.wred[par][wv][1];
      }
      const float E = 2.0f * es / nsum;                                  // model.py:319-325
      const float inc = __fadd_rn(__fmul_rn(E, dtf), sm.nz[buf][kk]);    // model.py:286
      X = __fadd_rn(X, inc);                                             // model.py:287
      const float s = inc / A;                                           // model.py:303
      const float rn = rsqrtf(nsum);   // lagged normalisation keeps |x| ~ 1
      float2 xp = make_float2(fmaf(s, y.x, a.x) * rn, fmaf(s, y.y, a.y) * rn);
      const float2 xn = cmul(sm.qs[buf][kk][i], xp);
      if (jq == 0) sm.xs[cur ^ 1][i] = xn;
      if (t == 0) sm.outs[kk] = A * X;                                   // model.py:251
      cur ^= 1;
      __syncthreads();
    }
    if (t < len) out[(size_t)b * L + k0 + t] = sm.outs[t];
    // outs is rewritten only after the next chunk's first two barriers
  }
  cp_async_wait<0>();
}

}  // namespace amps
