// RhoCMPS scan kernels (model.py:55-203): density-matrix evolution, one CTA per clip, one thread
// per matrix element.  In the interaction frame rho~ = P^dag rho P (P = diag(p_k)) the step is
//     rho~' = L rho~ L^dag,   L = N + s_k R,   N = I - (delta_t sigma^2/2) R^dag R
//     E     = Re tr((R + R^dag) rho~')          (model.py:189-196, on the un-normalised rho')
//     rho~_{k+1} = Q (rho~' / max(Re tr rho~', 1e-12)) Q^dag,   Q = diag(p_k conj(p_{k+1}))
// which is the reference's U rho U^dag with U = I + (-0.5 Rt^dag Rt delta_t sigma^2 + s Rt),
// Rt = P R P^dag (model.py:172-187).  Secondary path: correctness first, D <= 32.
#pragma once
#include "../../include/audiomps.h"
#include "amps_common.cuh"

namespace amps {

constexpr int RHO_MAX_D = 32;

inline size_t rho_workspace_bytes(int D, int B, int T) {
  (void)B;
  (void)T;
  return (D > 0 && D <= RHO_MAX_D) ? 256 : 0;
}

__global__ void fill_kernel(float* __restrict__ p, int n, float v) {
  const int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i < n) p[i] = v;
}

struct RhoArgs {
  const float2* R;      // [D][D]
  const float* freqs;   // [D]
  const float2* rho0;   // [D][D]
  const float* ttab;    // float32 time table
  int D;
  float A, dtf;
  double cprime;
  // data mode
  const float* x;       // [B][T]
  int T;
  float* loss;          // [B] or null
  // sample mode
  const float* noise;   // [L][n]
  int L, n;
  float* out;           // [n][L] or null
  float* purity;        // [n][L] or null
  // both
  float2* traj;         // [B][nsteps][D][D] lab frame, or null
};

template <bool SAMPLE>
__global__ void rho_scan_kernel(RhoArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = g.D, DD = D * D;
  float2* rho = reinterpret_cast<float2*>(smem_raw);  // [D][D] frame density matrix
  float2* Lm = rho + DD;                              // L = N + s R
  float2* Y = Lm + DD;                                // L rho
  float2* qv = Y + DD;                                // [D] q_k
  float2* pv = qv + D;                                // [D] p_{k+1}
  float* red = reinterpret_cast<float*>(pv + D);      // [32][2]

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const int b = blockIdx.x;
  const bool act = t < DD;
  const int a = act ? t / D : 0, c = act ? t % D : 0;
  const int nsteps = SAMPLE ? g.L : g.T - 1;

  // this thread's elements of R, N, S
  float2 Rab = make_float2(0.f, 0.f), Nab = Rab, Sba = Rab;
  if (act) {
    Rab = g.R[a * D + c];
    const float2 rba = g.R[c * D + a];
    // S_ba = R_ba + conj(R_ab)
    Sba = make_float2(rba.x + Rab.x, rba.y - Rab.y);
    double mr = 0.0, mi = 0.0;
    for (int m = 0; m < D; ++m) {
      const float2 u = g.R[m * D + a], v = g.R[m * D + c];
      mr += (double)u.x * v.x + (double)u.y * v.y;
      mi += (double)u.x * v.y - (double)u.y * v.x;
    }
    Nab = make_float2((float)((a == c ? 1.0 : 0.0) + g.cprime * mr), (float)(g.cprime * mi));
    rho[t] = g.rho0[t];
  }
  float X = 0.f;
  double lossacc = 0.0;
  __syncthreads();

  auto block_sum2 = [&](float v0, float v1, float& o0, float& o1) {
    v0 = warp_sum_f(v0);
    v1 = warp_sum_f(v1);
    if (lane == 0) {
      red[2 * warp] = v0;
      red[2 * warp + 1] = v1;
    }
    __syncthreads();
    float s0 = 0.f, s1 = 0.f;
    for (int wv = 0; wv < nw; ++wv) {
      s0 += red[2 * wv];
      s1 += red[2 * wv + 1];
    }
    o0 = s0;
    o1 = s1;
    __syncthreads();
  };

  for (int k = 0; k < nsteps; ++k) {
    // phases for this step (threads < D): q_k = p_k conj(p_{k+1}),  p_{k+1} for the lab frame
    if (t < D) {
      const float f = g.freqs[t];
      const double th0 = (double)__fmul_rn(f, g.ttab[k]);
      const double th1 = (double)__fmul_rn(f, g.ttab[k + 1]);
      double sn, cs;
      sincos(th0 - th1, &sn, &cs);
      qv[t] = make_float2((float)cs, (float)sn);
      if (g.traj) {
        sincos(th1, &sn, &cs);
        pv[t] = make_float2((float)cs, (float)sn);
      }
    }
    float inc;
    if (SAMPLE) {
      // E on the current normalised rho (model.py:162)
      float e = 0.f, dummy = 0.f;
      if (act) {
        const float2 r = rho[t];  // rho_ab, pairs with S_ba
        e = Sba.x * r.x - Sba.y * r.y;
      }
      float E, d2;
      block_sum2(e, dummy, E, d2);
      inc = __fadd_rn(__fmul_rn(E, g.dtf), g.noise[(size_t)k * g.n + b]);
      X = __fadd_rn(X, inc);
    } else {
      const float* xb = g.x + (size_t)b * g.T;
      inc = xb[k + 1] - xb[k];
    }
    const float s = inc / g.A;
    if (act) Lm[t] = make_float2(fmaf(s, Rab.x, Nab.x), fmaf(s, Rab.y, Nab.y));
    __syncthreads();
    if (act) {
      float2 acc = make_float2(0.f, 0.f);
      for (int m = 0; m < D; ++m) cmac(acc, Lm[a * D + m], rho[m * D + c]);
      Y[t] = acc;
    }
    __syncthreads();
    float2 rp = make_float2(0.f, 0.f);
    float e = 0.f, tr = 0.f;
    if (act) {
      for (int m = 0; m < D; ++m) cmac_cx(rp, Y[a * D + m], Lm[c * D + m]);
      e = Sba.x * rp.x - Sba.y * rp.y;  // Re(S_ba rho'_ab)
      tr = (a == c) ? rp.x : 0.f;
    }
    float E, TR;
    block_sum2(e, tr, E, TR);
    if (!SAMPLE && t == 0) lossacc -= log1p((double)((E * inc) / g.A));   // model.py:169-170
    const float inv = 1.0f / fmaxf(TR, 1e-12f);                            // model.py:198-203
    float2 rn = make_float2(0.f, 0.f);
    if (act) {
      rn = make_float2(rp.x * inv, rp.y * inv);
      // frame change: q_a conj(q_c)
      rn = cmul(qv[a], rn);
      rn = cmul_ca(qv[c], make_float2(rn.x, rn.y));
      // cmul_ca(q, z) = conj(q) * z
      rho[t] = rn;
    }
    __syncthreads();
    if (g.traj && act) {
      // lab frame: rho_ab = p_a rho~_ab conj(p_c)
      float2 lab = cmul(pv[a], rn);
      lab = cmul_ca(pv[c], lab);
      g.traj[((size_t)b * nsteps + k) * DD + t] = lab;
    }
    if (SAMPLE) {
      if (g.purity) {
        float pz = 0.f, dummy = 0.f;
        if (act) {
          const float2 r1 = rho[t], r2 = rho[c * D + a];
          pz = r1.x * r2.x - r1.y * r2.y;  // Re(rho_ab rho_ba)
        }
        float P, d2;
        block_sum2(pz, dummy, P, d2);
        if (t == 0) g.purity[(size_t)b * g.L + k] = P;
      }
      if (g.out && t == 0) g.out[(size_t)b * g.L + k] = g.A * X;          // model.py:112
    }
    __syncthreads();  // qv/pv/rho are rewritten at the top of the next step
  }
  if (!SAMPLE && g.loss && t == 0) g.loss[b] = (float)lossacc;
}

inline size_t rho_smem_bytes(int D) {
  return (size_t)(3 * D * D + 2 * D) * sizeof(float2) + (64 + 8) * sizeof(float);
}

inline int rho_block(int D) { return ((D * D + 31) / 32) * 32; }

inline int rho_launch_data(const amps_params* p, const float* ttab, const float* x, int B, int T,
                           float* loss, float2* traj, void* ws, cudaStream_t st) {
  (void)ws;
  RhoArgs g{};
  g.R = (const float2*)p->R_dev;
  g.freqs = p->freqs_dev;
  g.rho0 = (const float2*)p->rho0_dev;
  g.ttab = ttab;
  g.D = p->D;
  g.A = p->A;
  g.dtf = (float)p->delta_t;
  g.cprime = -p->delta_t * (double)p->sigma * (double)p->sigma / 2.0;
  g.x = x;
  g.T = T;
  g.loss = loss;
  g.traj = traj;
  rho_scan_kernel<false><<<B, rho_block(p->D), rho_smem_bytes(p->D), st>>>(g);
  return cudaGetLastError() == cudaSuccess ? 0 : AMPS_E_CUDA;
}

inline int rho_launch_sample(const amps_params* p, const float* ttab, const float* noise, int L,
                             int n, float* out, float2* traj, float* purity, void* ws,
                             cudaStream_t st) {
  (void)ws;
  RhoArgs g{};
  g.R = (const float2*)p->R_dev;
  g.freqs = p->freqs_dev;
  g.rho0 = (const float2*)p->rho0_dev;
  g.ttab = ttab;
  g.D = p->D;
  g.A = p->A;
  g.dtf = (float)p->delta_t;
  g.cprime = -p->delta_t * (double)p->sigma * (double)p->sigma / 2.0;
  g.noise = noise;
  g.L = L;
  g.n = n;
  g.out = out;
  g.purity = purity;
  g.traj = traj;
  rho_scan_kernel<true><<<n, rho_block(p->D), rho_smem_bytes(p->D), st>>>(g);
  return cudaGetLastError() == cudaSuccess ? 0 : AMPS_E_CUDA;
}

}  // namespace amps
