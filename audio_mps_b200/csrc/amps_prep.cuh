// Small parameter-only kernels: step-operator matrices, float32 time table, phase-rotation
// table, gradient finalisation, lab-frame trajectory.  All O(D^3) or O(T*D): off the hot path.
#pragma once
#include "amps_common.cuh"

namespace amps {

// t_0 = 0, t_{k+1} = fl32(t_k + dt32): the float32 running sum of model.py:16,157,281.
// Sequential by definition, but piecewise LINEAR in exact arithmetic: while t stays inside one
// binade (ulp u) and dt32/u is not a rounding tie, fl32(t + dt32) = t + rn(dt32/u)*u exactly.  One
// thread walks the <= ~40 binades (the steps that cross a binade boundary, the first steps and the
// one binade where dt32/u IS a tie are done with real float adds), then every thread fills the
// linear stretches in double (exact) -- so the table is rebuilt into the CALLER's workspace on every
// forward call in a few microseconds and no entry point depends on context-cached state.
struct TtSeg {
  int k0, m;     // out[k0 + j] = t0 + j*inc for j = 1..m
  float t0, inc;
};
constexpr int TT_MAXSEG = 96;

__host__ __device__ inline unsigned tt_bits(float v) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(v);
#else
  unsigned b;
  memcpy(&b, &v, 4);
  return b;
#endif
}
__host__ __device__ inline float tt_fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;   // IEEE single add, no contraction / excess precision
  return r;
#endif
}

// Writes the non-linear entries of out[0..n) directly and returns the linear stretches in segs.
__host__ __device__ inline int tt_build(float d, int n, float* out, TtSeg* segs, int maxseg) {
  int nseg = 0;
  if (n <= 0) return 0;
  float t = 0.f;
  out[0] = t;
  int k = 0;
  const unsigned db = tt_bits(d);
  const int dexp = (int)((db >> 23) & 0xff);
  const bool regular = d > 0.f && dexp > 0 && dexp < 255;       // positive normal step
  unsigned M = (db & 0x7fffffu) | 0x800000u;
  int tz = 0;
  while (tz < 24 && !((M >> tz) & 1u)) ++tz;
  while (k < n - 1) {
    const unsigned tb = tt_bits(t);
    const int texp = (int)((tb >> 23) & 0xff);
    const int j = texp - dexp;                                   // log2(ulp(t) / ulp(d))
    if (regular && t > 0.f && texp < 254 && j >= 1 && j != tz + 1 && nseg < maxseg) {
      // q = rn(d / u) as an integer (no tie here), a = (2^(E+1) - t) / u
      const unsigned long long q = j > 24 ? 0ull : (((unsigned long long)M + (1ull << (j - 1))) >> j);
      if (q == 0) {                                              // t no longer moves
        segs[nseg++] = TtSeg{k, n - 1 - k, t, 0.f};
        k = n - 1;
        break;
      }
      const unsigned long long a = (1ull << 24) - ((tb & 0x7fffffu) | 0x800000u);
      long long m = (long long)((a - 1) / q);                    // steps that stay below the binade top
      if (m > n - 1 - k) m = n - 1 - k;
      if (m > 0) {
        const double u = ldexp(1.0, texp - 127 - 23);
        const float inc = (float)((double)q * u);
        segs[nseg++] = TtSeg{k, (int)m, t, inc};
        t = (float)((double)t + (double)m * (double)inc);
        k += (int)m;
        if (k >= n - 1) break;
      }
    }
    t = tt_fadd(t, d);                                           // boundary / tie / start-up step
    out[++k] = t;
  }
  return nseg;
}

__global__ void prep_ttab_kernel(float dt32, int n, float* __restrict__ ttab) {
  __shared__ TtSeg segs[TT_MAXSEG];
  __shared__ int nseg;
  if (threadIdx.x == 0) nseg = tt_build(dt32, n, ttab, segs, TT_MAXSEG);
  __syncthreads();
  for (int s = 0; s < nseg; ++s) {
    const TtSeg g = segs[s];
    for (int j = 1 + (int)threadIdx.x; j <= g.m; j += (int)blockDim.x)
      ttab[g.k0 + j] = (float)((double)g.t0 + (double)j * (double)g.inc);
  }
}

// Padded [DP][DP] complex matrices from the effective R [D][D]:
//   matR = R, matRH = R^dag, matS = R + R^dag, matN = I + cprime * R^dag R  (exactly Hermitian)
__global__ void prep_mats_kernel(const float2* __restrict__ R, int D, int DP, double cprime,
                                 float2* __restrict__ matN, float2* __restrict__ matR,
                                 float2* __restrict__ matRH, float2* __restrict__ matS) {
  for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < DP * DP;
       idx += blockDim.x * gridDim.x) {
    const int i = idx / DP, j = idx % DP;
    const bool in = (i < D && j < D);
    const float2 rij = in ? R[i * D + j] : make_float2(0.f, 0.f);
    const float2 rji = in ? R[j * D + i] : make_float2(0.f, 0.f);
    matR[idx] = rij;
    matRH[idx] = make_float2(rji.x, -rji.y);
    matS[idx] = make_float2(rij.x + rji.x, rij.y - rji.y);
    if (i <= j) {
      double mr = 0.0, mi = 0.0;  // M_ij = sum_m conj(R_mi) R_mj
      if (in) {
        for (int m = 0; m < D; ++m) {
          const float2 a = R[m * D + i], bb = R[m * D + j];
          mr += (double)a.x * bb.x + (double)a.y * bb.y;
          mi += (double)a.x * bb.y - (double)a.y * bb.x;
        }
      }
      const double nr = (i == j ? 1.0 : 0.0) + cprime * mr;
      const double ni = (i == j ? 0.0 : cprime * mi);
      matN[i * DP + j] = make_float2((float)nr, (float)ni);
      if (i != j) matN[j * DP + i] = make_float2((float)nr, (float)(-ni));
    }
  }
}

// noise [L][n] (time-major, as tf.random_normal([length, num_samples]), model.py:246) -> [n][L]: the sampler
// kernels read one waveform's noise as a contiguous stream instead of one 4-byte element per 32-byte sector.
// 32 x 32 tiles through shared memory, coalesced on both sides.  grid = (ceil(n/32), ceil(L/32)), block = (32, 8).
__global__ void prep_transpose_kernel(const float* __restrict__ in, int L, int n, float* __restrict__ outT) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int l = l0 + r, c = n0 + threadIdx.x;
    if (l < L && c < n) tile[r][threadIdx.x] = in[(size_t)l * n + c];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = n0 + r, l = l0 + threadIdx.x;
    if (c < n && l < L) outT[(size_t)c * L + l] = tile[threadIdx.x][r];
  }
}

__global__ void prep_pad_vec_kernel(const float2* __restrict__ v, int D, int DP,
                                    float2* __restrict__ vp) {
  const int i = threadIdx.x + blockIdx.x * blockDim.x;
  if (i < DP) vp[i] = (i < D) ? v[i] : make_float2(0.f, 0.f);
}

// q_k[c] = p_k[c] conj(p_{k+1}[c]),  p_k[c] = exp(i * fl32(f_c * t_k))  (model.py:304-305), from the two
// float32 angles (exact in double), evaluated in double.
// The chain multiplies the state by q_k at EVERY step, so the accumulated phase is the PRODUCT of the
// ROUNDED table entries -- and their rounding errors are not random: fl32(f t_k) moves on a coarse grid
// (ulp ~1e-3 rad at 14000 rad), so q_k takes only two or three distinct values per component for
// thousands of steps and the same ~1.4e-8 rad rounding error repeats, i.e. accumulates LINEARLY
// (9e-4 rad over 64000 steps; it showed as 1.1e-3 in the full-length sampler golden, where E(psi) is a
// small residual of large terms and is fed back).  Fix: error feedback.  One thread walks QG
// consecutive steps of one component and keeps the double-precision product Q_j of the entries it has
// already rounded; entry j is  fl32( exp(i (theta_{k0} - theta_{k0+j+1})) / Q_j ),  so the product of
// the rounded entries tracks the exact rotation to ONE rounding (phase and modulus) anywhere inside a
// group, and the group boundaries add at most 1.5e-8 rad each (125 groups over 64000 steps).
// Measured (CPU emulation of the sampler arithmetic, 64000 steps): accumulated phase error 9e-4 ->
// 9e-7 rad, full-length sample error 1.1e-3 -> 2.6e-5 (the reference's own float32 run: 3.9e-5).
// Two phases per block (one group of QG steps x up to 32 components), software-pipelined over sub-tiles
// of QT steps: warps 1..7 evaluate the exact rotations exp(i (theta_first - theta_{j+1})) in double
// (independent, the expensive sincos) into shared memory; warp 0, one lane per component, runs the
// strictly sequential feedback recurrence over them (a few dependent FP64 FMAs per entry) and stores the
// rounded entries, 256 coalesced bytes per step.  ~35k cycles per group instead of ~800k for one thread
// doing both.
constexpr int QG = 512;    // steps per error-feedback group
constexpr int QT = 64;     // steps per pipeline sub-tile
constexpr int QC = 32;     // components per block (one lane of warp 0 each)
struct alignas(16) PhaseSmem {
  double2 ex[2][QT][QC];   // exact rotation from the group's first step to step j+1
};
// grid = (ceil(nsteps / QG), ceil(ld / QC)), block = 256.  Components c >= D (zero padding up to the row
// stride ld) get q = 1.  ptab (optional, Rho trajectories): p_{k+1} = exp(i fl32(f t_{k+1})).
__global__ void __launch_bounds__(256)
    prep_phase_tables_kernel(const float* __restrict__ freqs, int D, int ld, const float* __restrict__ ttab,
                             int nsteps, float2* __restrict__ qtab, float2* __restrict__ ptab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PhaseSmem& sm = *reinterpret_cast<PhaseSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int k0 = blockIdx.x * QG, n = min(QG, nsteps - k0);
  const int c0 = blockIdx.y * QC;
  const int ntiles = (n + QT - 1) / QT;
  const float* tt = ttab + k0;

  // exact rotations of sub-tile s, by the threads [first, first + nthr) of the block
  auto produce = [&](int s, int first, int nthr) {
    const int j0 = s * QT, len = min(QT, n - j0);
    for (int idx = t - first; idx < len * QC; idx += nthr) {
      const int j = idx / QC, cl = idx % QC, c = c0 + cl;
      double2 e = make_double2(1.0, 0.0);
      if (c < D) {
        const float f = freqs[c];
        const double th_first = (double)__fmul_rn(f, tt[0]);
        const double th = (double)__fmul_rn(f, tt[j0 + j + 1]);
        sincos(th_first - th, &e.y, &e.x);
        if (ptab) {
          double sn, cs;
          sincos(th, &sn, &cs);
          ptab[(size_t)(k0 + j0 + j) * ld + c] = make_float2((float)cs, (float)sn);
        }
      }
      sm.ex[s & 1][j][cl] = e;
    }
  };

  produce(0, 0, 256);
  __syncthreads();
  double Qx = 1.0, Qy = 0.0;            // product of the entries rounded so far (warp 0: one component per lane)
  for (int s = 0; s < ntiles; ++s) {
    if (warp == 0) {
      const int j0 = s * QT, len = min(QT, n - j0);
      const int c = c0 + lane;
      float2* dst = qtab + (size_t)(k0 + j0) * ld + c;
      for (int j = 0; j < len; ++j) {
        const double2 e = sm.ex[s & 1][j][lane];
        // entry = fl32(e / Q);  1/|Q|^2 = 2 - |Q|^2 to 1e-14 (|Q|^2 - 1 ~ 1e-7)
        const double inv = 2.0 - (Qx * Qx + Qy * Qy);
        const float2 q = make_float2((float)((e.x * Qx + e.y * Qy) * inv), (float)((e.y * Qx - e.x * Qy) * inv));
        const double nx = Qx * (double)q.x - Qy * (double)q.y, ny = Qx * (double)q.y + Qy * (double)q.x;
        Qx = nx;
        Qy = ny;
        if (c < ld) dst[(size_t)j * ld] = (c < D) ? q : make_float2(1.f, 0.f);
      }
    } else if (s + 1 < ntiles) {
      produce(s + 1, 32, 224);
    }
    __syncthreads();
  }
}

// Sum the per-clip partials over clips in a fixed order (deterministic).
//   Gtot[e] = sum_b G[b][e]  for e in [0, 3*DP*DP) ; gftot, lam0tot likewise.
__global__ void psi_reduce_clips_kernel(const float2* __restrict__ G, const float* __restrict__ gf,
                                        const float2* __restrict__ lam0, int B, int DP,
                                        float2* __restrict__ Gtot, float* __restrict__ gftot,
                                        float2* __restrict__ lam0tot, int lam_step = 1, int g_parts = 0) {
  // g_parts: number of partial tile sets in G when it differs from B (tensor-core tile kernel: B * nsplit)
  const int nG = 3 * DP * DP;
  const int total = nG + 2 * DP;
  const int nparts = g_parts > 0 ? g_parts : B;
  for (int e = threadIdx.x + blockIdx.x * blockDim.x; e < total; e += blockDim.x * gridDim.x) {
    if (e < nG) {
      double sx = 0.0, sy = 0.0;
      for (int b = 0; b < nparts; ++b) {
        const float2 v = G[(size_t)b * nG + e];
        sx += v.x;
        sy += v.y;
      }
      Gtot[e] = make_float2((float)sx, (float)sy);
    } else if (e < nG + DP) {
      const int c = e - nG;
      double s = 0.0;
      for (int b = 0; b < B; ++b) s += gf[(size_t)b * DP + c];
      gftot[c] = (float)s;
    } else {
      const int c = e - nG - DP;
      double sx = 0.0, sy = 0.0;
      for (int b = 0; b < B; b += lam_step) {   // virtual clips: only a clip's first chunk starts from psi_0
        const float2 v = lam0[(size_t)b * DP + c];
        sx += v.x;
        sy += v.y;
      }
      lam0tot[c] = make_float2((float)sx, (float)sy);
    }
  }
}

// Packed gradient wrt the effective parameters (grid = D CTAs, one row of gR each):
//   gR = G_R + G_E + cprime * R (G_N + G_N^dag)
//   gA = -(1/A) Re sum conj(G_R) R + sum_b gAdir[b]
//   out = [ gR (2 D^2) | gf (D) | gpsi0 (2 D) | gA | sum_b w_b loss_b ]
__global__ void psi_grad_finalize_kernel(const float2* __restrict__ Gtot,
                                         const float* __restrict__ gftot,
                                         const float2* __restrict__ lam0tot,
                                         const double* __restrict__ gAdir,
                                         const double* __restrict__ lossd,
                                         const float* __restrict__ w, int B,
                                         const float2* __restrict__ matR, int D, int DP,
                                         double cprime, AVal A_, float* __restrict__ out) {
  const float A = a_get(A_);
  __shared__ double red[32];
  const float2* GR = Gtot;
  const float2* GN = Gtot + DP * DP;
  const float2* GE = Gtot + 2 * DP * DP;
  // grid = D CTAs: CTA i writes row i of gR; CTA 0 also writes the vector / scalar slots
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const int idx = i * D + j;
    double cr = 0.0, ci = 0.0;  // (R H)_ij, H = GN + GN^dag
    for (int m = 0; m < D; ++m) {
      const float2 r = matR[i * DP + m];
      const float2 g1 = GN[m * DP + j], g2 = GN[j * DP + m];
      const double hx = (double)g1.x + g2.x, hy = (double)g1.y - g2.y;
      cr += r.x * hx - r.y * hy;
      ci += r.x * hy + r.y * hx;
    }
    const float2 gr = GR[i * DP + j], ge = GE[i * DP + j];
    out[2 * idx] = (float)((double)gr.x + ge.x + cprime * cr);
    out[2 * idx + 1] = (float)((double)gr.y + ge.y + cprime * ci);
  }
  if (blockIdx.x != 0) return;
  double part = 0.0;  // Re sum conj(GR) R
  for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
    const int a = idx / D, c = idx % D;
    const float2 gr = GR[a * DP + c], r = matR[a * DP + c];
    part += (double)gr.x * r.x + (double)gr.y * r.y;
  }
  float* o = out + 2 * D * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    o[c] = gftot[c];
    o[D + 2 * c] = lam0tot[c].x;
    o[D + 2 * c + 1] = lam0tot[c].y;
  }
  // block reduce `part`
  part = warp_sum_d(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) tot += red[wv];
    double ga = -tot / (double)A;
    double ls = 0.0;
    for (int b = 0; b < B; ++b) {
      ga += gAdir[b];
      ls += (double)w[b] * lossd[b];
    }
    o[3 * D] = (float)ga;
    o[3 * D + 1] = (float)ls;
  }
}

// Lab-frame normalised trajectory (model.py:231-240): psi_{k+1} = p_{k+1} x_{k+1} / |x_{k+1}|.
// One warp per (clip, step).
__global__ void psi_lab_traj_kernel(const float2* __restrict__ traj, const float* __restrict__ freqs,
                                    const float* __restrict__ ttab, int B, int T, int D, int DP,
                                    float2* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t total = (size_t)B * (T - 1);
  for (size_t item = wid; item < total; item += nw) {
    const int b = (int)(item / (T - 1)), k = (int)(item % (T - 1));
    const float2* xv = traj + ((size_t)b * T + k + 1) * DP;
    float n2 = 0.f;
    for (int c = lane; c < D; c += 32) n2 += cabs2(xv[c]);
    n2 = warp_sum_f(n2);
    const float rn = rsqrtf(fmaxf(n2, 1e-12f));          // model.py:331-332
    const float tk = ttab[k + 1];
    for (int c = lane; c < D; c += 32) {
      double sn, cs;
      sincos((double)__fmul_rn(freqs[c], tk), &sn, &cs);
      const float2 v = cmul(make_float2((float)cs, (float)sn), xv[c]);
      out[((size_t)b * (T - 1) + k) * D + c] = make_float2(v.x * rn, v.y * rn);
    }
  }
}

// -------------------------------------------------------------------------------------------
// a1 / a2 of the path in ONE launch each way: raw variables -> effective parameters and the
// regulariser (model.py:36-50, 218-222, 327-334; train.py:55-60), and the reverse chain.  Replaces
// ~60 tiny framework kernels per training step.  Single CTA; sums in double.
//   R = r_s (Rx + i Ry),  R_eff[i][j] = R[i][j] - R[j][j]  (the diagonal-broadcast quirk, model.py:42)
//   f_eff = f_s f_raw;   psi0 = z rsqrt(max(sum |z|^2, 1e-12)),  z = psi_x + i psi_y
//   reg = h_reg sum f_eff^2 + r_reg sum |R_eff|^2
// aux[0] = reg, aux[1] = rsqrt(max(|z|^2, eps)), aux[2] = 1 if the clamp is inactive
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
  for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) tot += red[wv];
  __syncthreads();
  return tot;
}

__global__ void psi_params_fwd_kernel(const float* __restrict__ Rx, const float* __restrict__ Ry,
                                      const float* __restrict__ fraw, const float* __restrict__ px,
                                      const float* __restrict__ py, int D, float r_s, float f_s, float h_reg,
                                      float r_reg, float2* __restrict__ R_eff, float* __restrict__ f_eff,
                                      float2* __restrict__ psi0, float* __restrict__ aux) {
  __shared__ double red[32];
  double accR = 0.0, accF = 0.0, accP = 0.0;
  for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
    const int j = idx % D;
    const float rx = __fsub_rn(__fmul_rn(r_s, Rx[idx]), __fmul_rn(r_s, Rx[j * D + j]));
    const float ry = __fsub_rn(__fmul_rn(r_s, Ry[idx]), __fmul_rn(r_s, Ry[j * D + j]));
    R_eff[idx] = make_float2(rx, ry);
    accR += (double)rx * rx + (double)ry * ry;
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float f = __fmul_rn(f_s, fraw[c]);
    f_eff[c] = f;
    accF += (double)f * f;
    const float n = hypotf(px[c], py[c]);      // |z| via complex abs, then squared (model.py:331-334)
    accP += (double)__fmul_rn(n, n);
  }
  accR = block_sum_d(accR, red);
  accF = block_sum_d(accF, red);
  accP = block_sum_d(accP, red);
  const float n2 = (float)accP;
  const float rs = rsqrtf(fmaxf(n2, 1e-12f));
  for (int c = threadIdx.x; c < D; c += blockDim.x) psi0[c] = make_float2(px[c] * rs, py[c] * rs);
  if (threadIdx.x == 0) {
    aux[0] = (float)((double)h_reg * accF + (double)r_reg * accR);
    aux[1] = rs;
    aux[2] = n2 > 1e-12f ? 1.f : 0.f;
  }
}

// gR, gpsi0: dL/dRe + i dL/dIm of the effective tensors; greg: dL/dreg (device scalar)
__global__ void psi_params_bwd_kernel(const float* __restrict__ Rx, const float* __restrict__ Ry,
                                      const float* __restrict__ fraw, const float* __restrict__ px,
                                      const float* __restrict__ py, int D, float r_s, float f_s, float h_reg,
                                      float r_reg, const float* __restrict__ aux, const float2* __restrict__ gR,
                                      const float* __restrict__ gf, const float2* __restrict__ gp,
                                      const float* __restrict__ greg, float* __restrict__ gRx,
                                      float* __restrict__ gRy, float* __restrict__ gfraw,
                                      float* __restrict__ gpx, float* __restrict__ gpy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* cs = reinterpret_cast<float2*>(smem_raw);      // [D] column sums of G
  __shared__ double red[32];
  const float gr = greg ? greg[0] : 0.f;
  const float kR = 2.0f * r_reg * gr;
  auto G = [&](int k, int j) {
    const float rx = __fsub_rn(__fmul_rn(r_s, Rx[k * D + j]), __fmul_rn(r_s, Rx[j * D + j]));
    const float ry = __fsub_rn(__fmul_rn(r_s, Ry[k * D + j]), __fmul_rn(r_s, Ry[j * D + j]));
    const float2 g = gR[k * D + j];
    return make_float2(fmaf(kR, rx, g.x), fmaf(kR, ry, g.y));
  };
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    double sx = 0.0, sy = 0.0;
    for (int k = 0; k < D; ++k) {
      const float2 g = G(k, j);
      sx += g.x;
      sy += g.y;
    }
    cs[j] = make_float2((float)sx, (float)sy);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
    const int i = idx / D, j = idx % D;
    float2 g = G(i, j);
    if (i == j) {                                  // R_eff[k][j] = R[k][j] - R[j][j] for every k
      g.x -= cs[j].x;
      g.y -= cs[j].y;
    }
    gRx[idx] = r_s * g.x;
    gRy[idx] = r_s * g.y;
  }
  const float rs = aux[1];
  const bool free_norm = aux[2] != 0.f;
  double dot = 0.0;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    gfraw[c] = f_s * fmaf(2.0f * h_reg * gr, __fmul_rn(f_s, fraw[c]), gf[c]);
    dot += (double)(px[c] * rs) * gp[c].x + (double)(py[c] * rs) * gp[c].y;   // Re <psi0, g>
  }
  dot = block_sum_d(dot, red);
  const float dt = free_norm ? (float)dot : 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    gpx[c] = (gp[c].x - (px[c] * rs) * dt) * rs;
    gpy[c] = (gp[c].y - (py[c] * rs) * dt) * rs;
  }
}

}  // namespace amps
