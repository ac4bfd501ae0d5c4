// PsiCMPS scan kernels (sm_100a): sequential persistent forward-loss kernel, adjoint backward,
// sampler.  One CTA owns one clip for the whole clip; the step operators live in registers,
// the state in shared memory, the waveform / phase table stream in through double-buffered
// cp.async.
//
// Formulation (validated against the op-for-op oracle, see DESIGN.md "Chain form"):
// in the interaction frame x_k = psi_k * conj(p_k) the reference step (model.py:276-334) is
//     x'_k    = N x_k + s_k R x_k,           N = I - (delta_t sigma^2 / 2) R^dag R,  s_k = inc_k / A
//     E_k     = x'_k^dag (R + R^dag) x'_k / |x_k|^2
//     loss   += -log(1 + (E_k inc_k) / A)
//     x_{k+1} = q_k * x'_k (* c),             q_k = p_k conj(p_{k+1})
// with p_k = exp(i fl32(f t_k)) and t_k the float32 running sum.  The state is carried
// UN-normalised (the loss is scale invariant) and rescaled by c once per chunk, so the only
// thing on the per-step critical path is one stacked [N;R] mat-vec; E_k, |x_k|^2, the log and
// the rescale are done lane-parallel over the CH steps of a chunk.
#pragma once
#include "amps_common.cuh"

namespace amps {

// -------------------------------------------------------------------------------------------
// shared-memory layouts
// -------------------------------------------------------------------------------------------
template <int DP>
struct alignas(16) FwdSmem {
  float2 xs[CH + 1][DP];   // x_{k0+kk}
  float2 xps[CH][DP];      // x'_{k0+kk}
  float2 qs[2][CH][DP];    // q_k, double buffered
  float es[CH][DP + 1];    // Re(conj(x'_i) (S x')_i)
  float ns[CH][DP + 1];    // |x_{k,i}|^2
  float wav[2][CH + 4];    // waveform samples k0..k0+len, double buffered
  float sv[CH];            // s_k
  float incv[CH];          // inc_k
  float scal[4];
};

template <int DP>
struct alignas(16) BwdSmem {
  float2 xs[2][CH + 1][DP];  // trajectory chunk, double buffered
  float2 qs[2][CH][DP];
  float2 xps[CH][DP];        // reconstructed x'_k
  float2 sps[CH][DP];        // S x'_k
  float2 mus[CH][DP];        // adjoint of x'_k
  float es[CH][DP + 1];
  float ns[CH][DP + 1];
  float wav[2][CH + 4];
  float tt[2][CH + 4];
  float sv[CH], incv[CH], dtk[CH], alphas[CH], betas[CH];
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
// P2: for every step of the chunk, (S x')_i and e_i = Re(conj(x'_i) (S x')_i)
// -------------------------------------------------------------------------------------------
template <int DP, int NQ, bool STORE_SP>
__device__ __forceinline__ void chunk_expectation(const float2 (&Sr)[DP / NQ],
                                                  const float2 (*xps)[DP], float2 (*sps)[DP],
                                                  float (*es)[DP + 1], int len, int i, int jq) {
  for (int kk = 0; kk < len; ++kk) {
    float2 part = matvec1<DP, NQ>(Sr, xps[kk], jq);
    part = group_sum<NQ>(part);
    if (STORE_SP && jq == 0) sps[kk][i] = part;
    if (jq == 1) {
      const float2 xpi = xps[kk][i];
      es[kk][i] = fmaf(xpi.x, part.x, xpi.y * part.y);
    }
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

  double lossacc = 0.0;
  if (nchunks > 0) issue_loads(0, 0);
  cp_async_commit();

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    const int k0 = c * CH;
    const int len = min(CH, nsteps - k0);
    if (c + 1 < nchunks) issue_loads(c + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    if (t < len) {
      const float inc = sm.wav[buf][t + 1] - sm.wav[buf][t];   // model.py:263
      sm.incv[t] = inc;
      sm.sv[t] = inc / A;                                       // model.py:303
    }
    if (t < DP) sm.ns[0][t] = cabs2(sm.xs[0][t]);
    __syncthreads();

    // ---- sequential chain: one stacked [N;R] mat-vec per step on the critical path --------
    for (int kk = 0; kk < len; ++kk) {
      float2 a, y;
      matvec2<DP, NQ>(Nr, Rr, sm.xs[kk], jq, a, y);
      a = group_sum<NQ>(a);
      y = group_sum<NQ>(y);
      const float s = sm.sv[kk];
      const float2 xp = make_float2(fmaf(s, y.x, a.x), fmaf(s, y.y, a.y));
      const float2 xn = cmul(sm.qs[buf][kk][i], xp);
      if (jq == 0) sm.xs[kk + 1][i] = xn;
      if (jq == 1) sm.xps[kk][i] = xp;
      if (jq == 2 && kk + 1 < CH) sm.ns[kk + 1][i] = cabs2(xn);
      __syncthreads();
    }

    // ---- lane-parallel part: E_k for every step of the chunk ------------------------------
    chunk_expectation<DP, NQ, false>(Sr, sm.xps, nullptr, sm.es, len, i, jq);
    __syncthreads();
    if (warp == 0) {
      const int kk = lane;
      if (kk < len) {
        float en = 0.f, nu2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < DP; ++r) {
          en += sm.es[kk][r];
          nu2 += sm.ns[kk][r];
        }
        const float E = en / nu2;                                // model.py:324-325 on x'
        const float z = (E * sm.incv[kk]) / A;                   // model.py:294
        lossacc -= log1p((double)z);
      }
      // rescale factor from |x_{k0+len}|^2
      float n2 = 0.f;
      for (int r = lane; r < DP; r += 32) n2 += cabs2(sm.xs[len][r]);
      n2 = warp_sum_f(n2);
      if (lane == 0) sm.scal[0] = rsqrtf(n2);
    }
    __syncthreads();
    const float sc = sm.scal[0];
    if (t < DP) {
      float2 v = sm.xs[len][t];
      v.x *= sc;
      v.y *= sc;
      sm.xs[len][t] = v;
    }
    if (t == 0 && scales) scales[(size_t)b * nchunks + c] = sc;
    __syncthreads();
    if (traj) {
      const float4* src = reinterpret_cast<const float4*>(&sm.xs[1][0]);
      float4* dst = reinterpret_cast<float4*>(traj + ((size_t)b * T + k0 + 1) * DP);
      for (int idx = t; idx < len * DP / 2; idx += NT) dst[idx] = src[idx];
      __syncthreads();
    }
    if (t < DP) sm.xs[0][t] = sm.xs[len][t];
    // the __syncthreads after the next cp.async wait orders this write before its readers
  }
  cp_async_wait<0>();

  if (warp == 0) {
    lossacc = warp_sum_d(lossacc);
    if (lane == 0) {
      loss[b] = (float)lossacc;
      if (lossd) lossd[b] = lossacc;
    }
  }
}

// -------------------------------------------------------------------------------------------
// K2: adjoint backward over the stored trajectory (replaces tf.gradients for train.py:89)
//   per-clip outputs: G[b][0]=sum_k s_k mu_k x_k^dag, G[b][1]=sum_k mu_k x_k^dag,
//                     G[b][2]=sum_k alpha_k x'_k x'_k^dag,  gf[b], lam0[b], gAdir[b]
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

  auto issue_loads = [&](int c, int buf) {
    const int k0 = c * CH;
    const int len = min(CH, nsteps - k0);
    const float2* xsrc = trb + (size_t)k0 * DP;
    float2* xdst = &sm.xs[buf][0][0];
    for (int idx = t; idx < (len + 1) * DP / 2; idx += NT) cp_async16(xdst + 2 * idx, xsrc + 2 * idx);
    const float2* qsrc = qtab + (size_t)k0 * DP;
    float2* qdst = &sm.qs[buf][0][0];
    for (int idx = t; idx < len * DP / 2; idx += NT) cp_async16(qdst + 2 * idx, qsrc + 2 * idx);
    for (int idx = t; idx <= len; idx += NT) {
      cp_async4(&sm.wav[buf][idx], xb + k0 + idx);
      cp_async4(&sm.tt[buf][idx], ttab + k0 + idx);
    }
  };

  float2 lam = make_float2(0.f, 0.f);  // adjoint of x_{k+1}, replicated over the NQ lanes
  float gf = 0.f;
  double gAacc = 0.0;

  if (nchunks > 0) issue_loads(nchunks - 1, (nchunks - 1) & 1);
  cp_async_commit();

  for (int c = nchunks - 1; c >= 0; --c) {
    const int buf = c & 1;
    const int k0 = c * CH;
    const int len = min(CH, nsteps - k0);
    if (c > 0) issue_loads(c - 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const float sc = scales[(size_t)b * nchunks + c];
    const float inv_sc = 1.0f / sc;

    if (t < len) {
      const float inc = sm.wav[buf][t + 1] - sm.wav[buf][t];
      sm.incv[t] = inc;
      sm.sv[t] = inc / A;
      sm.dtk[t] = sm.tt[buf][t + 1] - sm.tt[buf][t];
    }
    // P1: x'_k = conj(q_k) x_{k+1} / c_k ; |x_k|^2
    for (int idx = t; idx < len * DP; idx += NT) {
      const int kk = idx / DP, r = idx % DP;
      float2 xp = cmul_ca(sm.qs[buf][kk][r], sm.xs[buf][kk + 1][r]);
      if (kk == len - 1) {
        xp.x *= inv_sc;
        xp.y *= inv_sc;
      }
      sm.xps[kk][r] = xp;
      sm.ns[kk][r] = cabs2(sm.xs[buf][kk][r]);
    }
    __syncthreads();
    // P2: S x' and e_i
    chunk_expectation<DP, NQ, true>(Sr, sm.xps, sm.sps, sm.es, len, i, jq);
    __syncthreads();
    // P3: per-step scalars
    if (warp == 0) {
      const int kk = lane;
      if (kk < len) {
        float en = 0.f, nu2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < DP; ++r) {
          en += sm.es[kk][r];
          nu2 += sm.ns[kk][r];
        }
        const float E = en / nu2;
        const float inc = sm.incv[kk];
        const float arg = 1.0f + (E * inc) / A;
        const float gE = wb * (-sm.sv[kk] / arg);
        const float alpha = 2.0f * gE / nu2;
        sm.alphas[kk] = alpha;
        sm.betas[kk] = -alpha * E;
        gAacc += (double)wb * (double)E * (double)inc / ((double)A * (double)A * (double)arg);
      }
    }
    __syncthreads();

    // ---- sequential adjoint chain ---------------------------------------------------------
    for (int kk = len - 1; kk >= 0; --kk) {
      const float2 q = sm.qs[buf][kk][i];
      const float2 xn = sm.xs[buf][kk + 1][i];
      gf = fmaf(sm.dtk[kk], lam.x * xn.y - lam.y * xn.x, gf);   // Im(conj(lam) x_{k+1})
      float2 mu = cmul_ca(q, lam);
      if (kk == len - 1) {
        mu.x *= sc;
        mu.y *= sc;
      }
      const float al = sm.alphas[kk];
      const float2 sp = sm.sps[kk][i];
      mu.x = fmaf(al, sp.x, mu.x);
      mu.y = fmaf(al, sp.y, mu.y);
      if (jq == 0) sm.mus[kk][i] = mu;
      __syncthreads();
      float2 a, h;
      matvec2<DP, NQ>(Nr, Hr, sm.mus[kk], jq, a, h);
      a = group_sum<NQ>(a);
      h = group_sum<NQ>(h);
      const float s = sm.sv[kk];
      const float be = sm.betas[kk];
      const float2 xk = sm.xs[buf][kk][i];
      lam.x = fmaf(be, xk.x, fmaf(s, h.x, a.x));
      lam.y = fmaf(be, xk.y, fmaf(s, h.y, a.y));
    }

    // ---- parameter-gradient tiles (rank-1 updates, lane-parallel over the chunk) ----------
    for (int kk = 0; kk < len; ++kk) {
      const float2 mui = sm.mus[kk][i];
      const float2 xpi = sm.xps[kk][i];
      const float s = sm.sv[kk];
      const float al = sm.alphas[kk];
      const float2 u1 = make_float2(s * mui.x, s * mui.y);
      const float2 u3 = make_float2(al * xpi.x, al * xpi.y);
#pragma unroll
      for (int m = 0; m < CPT / 2; ++m) {
        const float4 xv = *reinterpret_cast<const float4*>(&sm.xs[buf][kk][2 * NQ * m + 2 * jq]);
        const float4 pv = *reinterpret_cast<const float4*>(&sm.xps[kk][2 * NQ * m + 2 * jq]);
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
    __syncthreads();
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
  if (warp == 0) {
    gAacc = warp_sum_d(gAacc);
    if (lane == 0) gAdir[b] = gAacc;
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
