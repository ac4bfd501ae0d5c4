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
#include "amps_prep.cuh"

namespace amps {

constexpr int RHO_MAX_D = 32;

struct RhoWs {
  size_t ttab, qtab, ptab, ftraj, G, acc, lam0, gAdir, lossd, scratch, total;
};
inline RhoWs rho_ws_layout(int D, int B, int T, bool save) {
  RhoWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t DD = (size_t)D * D;
  w.lossd = take((size_t)(B > 0 ? B : 1) * sizeof(double));
  // tables of one call (float32 t_k, q_k, p_{k+1}): in the caller's workspace, rebuilt by every forward
  w.ttab = take((size_t)(T + 2) * sizeof(float));
  w.qtab = take((size_t)(T > 0 ? T : 1) * D * sizeof(float2));
  w.ptab = take((size_t)(T > 0 ? T : 1) * D * sizeof(float2));
  if (save) {
    w.ftraj = take((size_t)B * T * DD * sizeof(float2));
    w.G = take((size_t)B * 3 * DD * sizeof(float2));
    w.acc = take((size_t)B * DD * sizeof(float));
    w.lam0 = take((size_t)B * DD * sizeof(float2));
    w.gAdir = take((size_t)B * sizeof(double));
    w.scratch = take(4 * DD * sizeof(float2));
  }
  w.total = off;
  return w;
}
inline size_t rho_workspace_bytes(int D, int B, int T, bool save = false) {
  return (D > 0 && D <= RHO_MAX_D && B >= 0 && T >= 0) ? rho_ws_layout(D, B, T, save).total : 0;
}

__global__ void fill_kernel(float* __restrict__ p, int n, float v) {
  const int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i < n) p[i] = v;
}

// Phase tables of a call (q_k, and p_{k+1} for lab-frame trajectories): prep_phase_tables_kernel, amps_prep.cuh.

struct RhoArgs {
  const float2* R;      // [D][D]
  const float* freqs;   // [D]
  const float2* rho0;   // [D][D]
  const float* ttab;    // float32 time table
  const float2* qtab;   // [nsteps][D] q_k = p_k conj(p_{k+1})   (rho_prep_phase_kernel)
  const float2* ptab;   // [nsteps][D] p_{k+1}, only when a lab-frame trajectory is requested
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
  float2* ftraj;        // [B][T][D][D] interaction-frame rho at the START of every step (for the adjoint), or null
  double* lossd;        // [B] or null
};

template <bool SAMPLE>
__global__ void __launch_bounds__(RHO_MAX_D * RHO_MAX_D) rho_scan_kernel(RhoArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // shared matrices have row stride LD = D + 1: column-indexed reads (X[c * LD + m], consecutive c per
  // lane) then hit distinct banks instead of a D-way conflict
  const int D = g.D, DD = D * D, LD = D + 1, DL = D * LD;
  float2* rho = reinterpret_cast<float2*>(smem_raw);  // [D][D] frame density matrix
  float2* Lm = rho + DL;                              // L = N + s R
  float2* Y = Lm + DL;                                // L rho
  float2* qv = Y + DL;                                // [D] q_k
  float2* pv = qv + D;                                // [D] p_{k+1}
  float* red = reinterpret_cast<float*>(pv + D);      // [32][2]

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const int b = blockIdx.x;
  const bool act = t < DD;
  const int a = act ? t / D : 0, c = act ? t % D : 0, ts = a * LD + c;
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
    rho[ts] = g.rho0[t];
    if (!SAMPLE && g.ftraj) g.ftraj[(size_t)b * g.T * DD + t] = g.rho0[t];
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

  // The per-step global inputs (waveform / noise sample, phases) are fetched ONE STEP AHEAD into
  // registers, so their L2 latency is off the step's critical path.
  const float* xb = SAMPLE ? nullptr : g.x + (size_t)b * g.T;
  // (raw loads only: the increment x[k+1] - x[k] is formed when it is used, one step later, so that no
  // instruction of the current step waits on the prefetch)
  auto raw_at = [&](int k) -> float {          // noise[k] (sampling) or waveform sample x[k + 1] (data)
    if (k >= nsteps) return 0.f;
    return SAMPLE ? g.noise[(size_t)k * g.n + b] : xb[k + 1];
  };
  float x_lo = (SAMPLE || nsteps == 0) ? 0.f : xb[0];
  float raw_next = raw_at(0);
  float2 q_next = make_float2(1.f, 0.f), p_next = q_next;
  if (t < D && nsteps > 0) {
    q_next = g.qtab[t];
    if (g.traj) p_next = g.ptab[t];
  }
  for (int k = 0; k < nsteps; ++k) {
    // phases for this step (threads < D): q_k = p_k conj(p_{k+1}),  p_{k+1} for the lab frame
    const float raw_cur = raw_next;
    raw_next = raw_at(k + 1);
    const float in_cur = SAMPLE ? raw_cur : raw_cur - x_lo;                // model.py:134-135
    x_lo = raw_cur;
    if (t < D) {
      qv[t] = q_next;
      if (g.traj) pv[t] = p_next;
      if (k + 1 < nsteps) {
        q_next = g.qtab[(size_t)(k + 1) * D + t];
        if (g.traj) p_next = g.ptab[(size_t)(k + 1) * D + t];
      }
    }
    float inc;
    if (SAMPLE) {
      // E on the current normalised rho (model.py:162)
      float e = 0.f, dummy = 0.f;
      if (act) {
        const float2 r = rho[ts];  // rho_ab, pairs with S_ba
        e = Sba.x * r.x - Sba.y * r.y;
      }
      float E, d2;
      block_sum2(e, dummy, E, d2);
      inc = __fadd_rn(__fmul_rn(E, g.dtf), in_cur);
      X = __fadd_rn(X, inc);
    } else {
      inc = in_cur;
    }
    const float s = inc / g.A;
    if (act) Lm[ts] = make_float2(fmaf(s, Rab.x, Nab.x), fmaf(s, Rab.y, Nab.y));
    __syncthreads();
    if (act) {
      float2 acc = make_float2(0.f, 0.f);
      for (int m = 0; m < D; ++m) cmac(acc, Lm[a * LD + m], rho[m * LD + c]);
      Y[ts] = acc;
    }
    __syncthreads();
    float2 rp = make_float2(0.f, 0.f);
    float e = 0.f, tr = 0.f;
    if (act) {
      for (int m = 0; m < D; ++m) cmac_cx(rp, Y[a * LD + m], Lm[c * LD + m]);
      e = Sba.x * rp.x - Sba.y * rp.y;  // Re(S_ba rho'_ab)
      tr = (a == c) ? rp.x : 0.f;
    }
    float E, TR;
    block_sum2(e, tr, E, TR);
    if (!SAMPLE && t == 0) lossacc -= (double)log1pf((E * inc) / g.A);    // model.py:169-170 (fp32 log, fp64 sum)
    const float inv = 1.0f / fmaxf(TR, 1e-12f);                            // model.py:198-203
    float2 rn = make_float2(0.f, 0.f);
    if (act) {
      rn = make_float2(rp.x * inv, rp.y * inv);
      // frame change: q_a conj(q_c)
      rn = cmul(qv[a], rn);
      rn = cmul_ca(qv[c], make_float2(rn.x, rn.y));
      // cmul_ca(q, z) = conj(q) * z
      rho[ts] = rn;
      if (!SAMPLE && g.ftraj) g.ftraj[((size_t)b * g.T + k + 1) * DD + t] = rn;
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
          const float2 r1 = rho[ts], r2 = rho[c * LD + a];
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
  if (!SAMPLE && g.lossd && t == 0) g.lossd[b] = lossacc;
}

// -------------------------------------------------------------------------------------------
// adjoint of the rho loss fold (DESIGN.md "Rho adjoint"), one CTA per clip, one thread per element.
// k descending, Lam = adjoint of rho~_{k+1}:
//   Grp   = conj(q_a) q_c Lam_ac                       (through the frame change)
//   GP    = Grp/tau - Re tr(Grp^dag P)/tau^2 I + gE S  (through the trace normalisation and the loss term)
//   Lam   = L^dag GP L,    GL = GP (L rho^dag) + GP^dag (L rho)
// per-clip outputs: sum_k s_k GL, sum_k GL, sum_k gE (P + P^dag), freq accumulator, Lam_0, gA(direct)
// -------------------------------------------------------------------------------------------
struct RhoBwdArgs {
  const float2* R;
  const float* freqs;
  const float* ttab;
  const float2* qtab;   // [nsteps][D]
  const float* x;       // [B][T]
  const float* w;       // [B]
  const float2* ftraj;  // [B][T][D][D]
  int D, T;
  float A;
  double cprime;
  float2* Gout;         // [B][3][D][D]  (s GL, GL, gE(P+P^dag))
  float* accout;        // [B][D][D]
  float2* lam0out;      // [B][D][D]
  double* gAdir;        // [B]
};

// launch bounds: D = 32 runs 1024 threads, i.e. at most 64 registers per thread
__global__ void __launch_bounds__(RHO_MAX_D * RHO_MAX_D) rho_bwd_kernel(RhoBwdArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = g.D, DD = D * D, LD = D + 1, DL = D * LD;   // padded row stride, as in rho_scan_kernel
  float2* rho = reinterpret_cast<float2*>(smem_raw);
  float2* Lm = rho + DL;
  float2* Y = Lm + DL;      // L rho
  float2* Y2 = Y + DL;      // L rho^dag
  float2* Ps = Y2 + DL;     // P
  float2* GPs = Ps + DL;    // adjoint of P
  float2* Z1 = GPs + DL;    // GP L
  float2* qv = Z1 + DL;     // [D]
  float* red = reinterpret_cast<float*>(qv + D);   // [32][3]

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const int b = blockIdx.x;
  const bool act = t < DD;
  const int a = act ? t / D : 0, c = act ? t % D : 0, ts = a * LD + c;
  const int nsteps = g.T - 1;
  const float wb = g.w[b];
  const float* xb = g.x + (size_t)b * g.T;
  const float2* fb = g.ftraj + (size_t)b * g.T * DD;

  float2 Rac = make_float2(0.f, 0.f), Nac = Rac, Sac = Rac, Sca = Rac;
  if (act) {
    Rac = g.R[a * D + c];
    const float2 rca = g.R[c * D + a];
    Sac = make_float2(Rac.x + rca.x, Rac.y - rca.y);   // S = R + R^dag
    Sca = make_float2(Sac.x, -Sac.y);
    double mr = 0.0, mi = 0.0;
    for (int m = 0; m < D; ++m) {
      const float2 u = g.R[m * D + a], v = g.R[m * D + c];
      mr += (double)u.x * v.x + (double)u.y * v.y;
      mi += (double)u.x * v.y - (double)u.y * v.x;
    }
    Nac = make_float2((float)((a == c ? 1.0 : 0.0) + g.cprime * mr), (float)(g.cprime * mi));
  }
  float2 Lam = make_float2(0.f, 0.f), GLs = Lam, GL1 = Lam, GPacc = Lam;
  float acc = 0.f;
  double gAacc = 0.0;

  auto block_sum3 = [&](float v0, float v1, float v2, float& o0, float& o1, float& o2) {
    v0 = warp_sum_f(v0);
    v1 = warp_sum_f(v1);
    v2 = warp_sum_f(v2);
    if (lane == 0) {
      red[3 * warp] = v0;
      red[3 * warp + 1] = v1;
      red[3 * warp + 2] = v2;
    }
    __syncthreads();
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int wv = 0; wv < nw; ++wv) {
      s0 += red[3 * wv];
      s1 += red[3 * wv + 1];
      s2 += red[3 * wv + 2];
    }
    o0 = s0;
    o1 = s1;
    o2 = s2;
    __syncthreads();
  };

  // inputs of step k-1 are fetched during step k (registers): waveform increment, time difference,
  // phases and the stored rho~_{k-1}; rho~_{k+1} of a step is the previous step's rho~_k.
  // raw loads only (x[k], t_k of the NEXT step); differences are formed when used
  float x_hi = nsteps > 0 ? xb[nsteps] : 0.f, t_hi = nsteps > 0 ? g.ttab[nsteps] : 0.f;
  float x_next = nsteps > 0 ? xb[nsteps - 1] : 0.f, t_next = nsteps > 0 ? g.ttab[nsteps - 1] : 0.f;
  float2 q_next = make_float2(1.f, 0.f), rho_next = make_float2(0.f, 0.f), rho_prev = rho_next;
  if (nsteps > 0) {
    if (t < D) q_next = g.qtab[(size_t)(nsteps - 1) * D + t];
    if (act) {
      rho_next = fb[(size_t)(nsteps - 1) * DD + t];
      rho_prev = fb[(size_t)nsteps * DD + t];
    }
  }
  for (int k = nsteps - 1; k >= 0; --k) {
    const float x_k = x_next, t_k = t_next;
    const float inc = x_hi - x_k, dl = t_k - t_hi;       // x[k+1] - x[k],  t_k - t_{k+1}
    x_hi = x_k;
    t_hi = t_k;
    if (k > 0) {
      x_next = xb[k - 1];
      t_next = g.ttab[k - 1];
    }
    const float s = inc / g.A;
    const float2 rk = rho_next, rn = rho_prev;
    if (act) {
      rho[ts] = rk;
      Lm[ts] = make_float2(fmaf(s, Rac.x, Nac.x), fmaf(s, Rac.y, Nac.y));
      if (k > 0) rho_next = fb[(size_t)(k - 1) * DD + t];
    }
    rho_prev = rk;
    if (t < D) {
      qv[t] = q_next;
      if (k > 0) q_next = g.qtab[(size_t)(k - 1) * D + t];
    }
    __syncthreads();
    acc = fmaf(dl, Lam.x * rn.y - Lam.y * rn.x, acc);     // dl Im(conj(Lam) rho~_{k+1})
    if (act) {
      float2 y = make_float2(0.f, 0.f), y2 = y;
      for (int m = 0; m < D; ++m) {
        cmac(y, Lm[a * LD + m], rho[m * LD + c]);
        cmac_cx(y2, Lm[a * LD + m], rho[c * LD + m]);
      }
      Y[ts] = y;
      Y2[ts] = y2;
    }
    __syncthreads();
    float2 p = make_float2(0.f, 0.f), grp = p;
    float e = 0.f, tr = 0.f, dot = 0.f;
    if (act) {
      for (int m = 0; m < D; ++m) cmac_cx(p, Y[a * LD + m], Lm[c * LD + m]);
      Ps[ts] = p;
      e = Sca.x * p.x - Sca.y * p.y;                      // Re(S_ca P_ac)
      tr = (a == c) ? p.x : 0.f;
      grp = cmul(cmul_ca(qv[a], Lam), qv[c]);             // conj(q_a) q_c Lam
      dot = grp.x * p.x + grp.y * p.y;                    // Re(conj(Grp) P)
    }
    float E, TAU, DOT;
    block_sum3(e, tr, dot, E, TAU, DOT);
    const float arg = 1.0f + (E * inc) / g.A;
    const float gE = wb * (-s / arg);
    if (t == 0) gAacc += (double)(wb * E * inc / (g.A * g.A * arg));
    if (act) {
      const float it = 1.0f / TAU;
      float2 gp = make_float2(grp.x * it + gE * Sac.x, grp.y * it + gE * Sac.y);
      if (a == c) gp.x -= DOT * it * it;
      GPs[ts] = gp;
      const float2 pca = Ps[c * LD + a];
      GPacc.x = fmaf(gE, p.x + pca.x, GPacc.x);
      GPacc.y = fmaf(gE, p.y - pca.y, GPacc.y);
    }
    __syncthreads();
    if (act) {
      float2 z = make_float2(0.f, 0.f), gl = z;
      for (int m = 0; m < D; ++m) {
        cmac(z, GPs[a * LD + m], Lm[m * LD + c]);
        cmac(gl, GPs[a * LD + m], Y2[m * LD + c]);
        // GP^dag Y : conj(GP[m][a]) Y[m][c]
        const float2 gm = GPs[m * LD + a], ym = Y[m * LD + c];
        gl.x = fmaf(gm.x, ym.x, fmaf(gm.y, ym.y, gl.x));
        gl.y = fmaf(gm.x, ym.y, fmaf(-gm.y, ym.x, gl.y));
      }
      Z1[ts] = z;
      GLs.x = fmaf(s, gl.x, GLs.x);
      GLs.y = fmaf(s, gl.y, GLs.y);
      GL1.x += gl.x;
      GL1.y += gl.y;
    }
    __syncthreads();
    if (act) {
      float2 ln = make_float2(0.f, 0.f);
      for (int m = 0; m < D; ++m) {                       // L^dag Z1 : conj(L[m][a]) Z1[m][c]
        const float2 lm = Lm[m * LD + a], zm = Z1[m * LD + c];
        ln.x = fmaf(lm.x, zm.x, fmaf(lm.y, zm.y, ln.x));
        ln.y = fmaf(lm.x, zm.y, fmaf(-lm.y, zm.x, ln.y));
      }
      Lam = ln;
    }
    __syncthreads();
  }
  if (act) {
    float2* Gb = g.Gout + (size_t)b * 3 * DD;
    Gb[t] = GLs;
    Gb[DD + t] = GL1;
    Gb[2 * DD + t] = GPacc;
    g.accout[(size_t)b * DD + t] = acc;
    g.lam0out[(size_t)b * DD + t] = Lam;
  }
  if (t == 0) g.gAdir[b] = gAacc;
}

// packed gradient: [ gR (2 D^2) | gf (D) | grho0 (2 D^2) | gA | sum_b w_b loss_b ]   (single CTA)
__global__ void rho_grad_finalize_kernel(const float2* __restrict__ G, const float* __restrict__ accs,
                                         const float2* __restrict__ lam0, const double* __restrict__ gAdir,
                                         const double* __restrict__ lossd, const float* __restrict__ w, int B,
                                         const float2* __restrict__ R, int D, double cprime, float A,
                                         float2* __restrict__ scratch, float* __restrict__ out) {
  __shared__ double redd[32];
  const int DD = D * D;
  // scratch: [0] sum s GL, [1] sum GL, [2] sum gE(P+P^dag), then acc totals as floats
  float* acct = reinterpret_cast<float*>(scratch + 3 * DD);
  for (int e = threadIdx.x; e < 3 * DD; e += blockDim.x) {
    double sx = 0.0, sy = 0.0;
    for (int b = 0; b < B; ++b) {
      const float2 v = G[(size_t)b * 3 * DD + e];
      sx += v.x;
      sy += v.y;
    }
    scratch[e] = make_float2((float)sx, (float)sy);
  }
  for (int e = threadIdx.x; e < DD; e += blockDim.x) {
    double s = 0.0, lx = 0.0, ly = 0.0;
    for (int b = 0; b < B; ++b) {
      s += accs[(size_t)b * DD + e];
      lx += lam0[(size_t)b * DD + e].x;
      ly += lam0[(size_t)b * DD + e].y;
    }
    acct[e] = (float)s;
    out[2 * DD + D + 2 * e] = (float)lx;       // grho0
    out[2 * DD + D + 2 * e + 1] = (float)ly;
  }
  __syncthreads();
  const float2* GLs = scratch;
  const float2* GL1 = scratch + DD;
  const float2* GP = scratch + 2 * DD;
  double part = 0.0;
  for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
    const int i = idx / D, j = idx % D;
    double cr = 0.0, ci = 0.0;   // (R H)_ij, H = GL1 + GL1^dag
    for (int m = 0; m < D; ++m) {
      const float2 r = R[i * D + m];
      const float2 g1 = GL1[m * D + j], g2 = GL1[j * D + m];
      const double hx = (double)g1.x + g2.x, hy = (double)g1.y - g2.y;
      cr += r.x * hx - r.y * hy;
      ci += r.x * hy + r.y * hx;
    }
    out[2 * idx] = (float)((double)GLs[idx].x + GP[idx].x + cprime * cr);
    out[2 * idx + 1] = (float)((double)GLs[idx].y + GP[idx].y + cprime * ci);
    const float2 r = R[idx];
    part += (double)GLs[idx].x * r.x + (double)GLs[idx].y * r.y;
  }
  for (int cidx = threadIdx.x; cidx < D; cidx += blockDim.x) {   // gf[c] = sum_a acc[a][c] - sum_b acc[c][b]
    double s = 0.0;
    for (int m = 0; m < D; ++m) s += (double)acct[m * D + cidx] - (double)acct[cidx * D + m];
    out[2 * DD + cidx] = (float)s;
  }
  part = warp_sum_d(part);
  if ((threadIdx.x & 31) == 0) redd[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) tot += redd[wv];
    double ga = -tot / (double)A, ls = 0.0;
    for (int b = 0; b < B; ++b) {
      ga += gAdir[b];
      ls += (double)w[b] * lossd[b];
    }
    out[4 * DD + D] = (float)ga;
    out[4 * DD + D + 1] = (float)ls;
  }
}

inline size_t rho_smem_bytes(int D) {
  return (size_t)(3 * D * (D + 1) + 2 * D) * sizeof(float2) + (64 + 8) * sizeof(float);
}

inline int rho_block(int D) { return ((D * D + 31) / 32) * 32; }

inline size_t rho_bwd_smem_bytes(int D) {
  return (size_t)(7 * D * (D + 1) + D) * sizeof(float2) + 96 * sizeof(float);
}

inline int rho_launch_bwd(const amps_params* p, const float* ttab, const float2* qtab, const float* x, int B, int T,
                          const float* w, char* ws, const RhoWs& L, float* grad, cudaStream_t st) {
  RhoBwdArgs g{};
  g.qtab = qtab;
  g.R = (const float2*)p->R_dev;
  g.freqs = p->freqs_dev;
  g.ttab = ttab;
  g.x = x;
  g.w = w;
  g.ftraj = (const float2*)(ws + L.ftraj);
  g.D = p->D;
  g.T = T;
  g.A = p->A;
  g.cprime = -p->delta_t * (double)p->sigma * (double)p->sigma / 2.0;
  g.Gout = (float2*)(ws + L.G);
  g.accout = (float*)(ws + L.acc);
  g.lam0out = (float2*)(ws + L.lam0);
  g.gAdir = (double*)(ws + L.gAdir);
  const size_t smem = rho_bwd_smem_bytes(p->D);
  rho_bwd_kernel<<<B, rho_block(p->D), smem, st>>>(g);
  if (cudaGetLastError() != cudaSuccess) return AMPS_E_CUDA;
  rho_grad_finalize_kernel<<<1, 256, 0, st>>>((const float2*)(ws + L.G), (const float*)(ws + L.acc),
                                              (const float2*)(ws + L.lam0), (const double*)(ws + L.gAdir),
                                              (const double*)(ws + L.lossd), w, B, (const float2*)p->R_dev, p->D,
                                              g.cprime, p->A, (float2*)(ws + L.scratch), grad);
  return cudaGetLastError() == cudaSuccess ? 0 : AMPS_E_CUDA;
}

inline int rho_launch_data(const amps_params* p, const float* ttab, const float2* qtab, const float2* ptab,
                           const float* x, int B, int T,
                           float* loss, float2* traj, float2* ftraj, double* lossd, cudaStream_t st) {
  RhoArgs g{};
  g.qtab = qtab;
  g.ptab = ptab;
  g.ftraj = ftraj;
  g.lossd = lossd;
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

inline int rho_launch_sample(const amps_params* p, const float* ttab, const float2* qtab, const float2* ptab,
                             const float* noise, int L,
                             int n, float* out, float2* traj, float* purity, void* ws,
                             cudaStream_t st) {
  (void)ws;
  RhoArgs g{};
  g.qtab = qtab;
  g.ptab = ptab;
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
