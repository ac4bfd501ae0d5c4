/* Precision-generic body of the C restatement; included twice by cmps_ref.c with
 *   REAL = float  / SUF = f32   (the reference's arithmetic: float32 / complex64)
 *   REAL = double / SUF = f64   (exact-arithmetic value of the same function)
 * TEST INFRASTRUCTURE ONLY (see cmps_ref.c).  Follows /root/reference/model.py line by line in the
 * LAB frame with a normalisation every step -- deliberately NOT the interaction-frame chain form
 * the CUDA kernels use, so that the two are independent derivations.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

typedef struct { REAL re, im; } FN(cplx);
#define CX FN(cplx)

static inline CX FN(cmul)(CX a, CX b) { CX r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline CX FN(cconj)(CX a) { CX r = {a.re, -a.im}; return r; }
static inline CX FN(cadd)(CX a, CX b) { CX r = {a.re + b.re, a.im + b.im}; return r; }
static inline CX FN(cscale)(CX a, REAL s) { CX r = {a.re * s, a.im * s}; return r; }

/* phases p_c = exp(i * fl32(f_c * t)), model.py:304-305 (angle is the float32 product in both modes) */
static void FN(phases)(int D, const float* f, float t32, CX* p) {
  for (int c = 0; c < D; ++c) {
    float ang = f[c] * t32;
#if IS_F32
    p[c].re = cosf(ang); p[c].im = sinf(ang);
#else
    p[c].re = cos((double)ang); p[c].im = sin((double)ang);
#endif
  }
}

/* y = M x (complex DxD row-major times vector) ; einsum 'bc,ac->ab' for one clip, model.py:309-310 */
static void FN(matvec)(int D, const CX* M, const CX* x, CX* y) {
  for (int i = 0; i < D; ++i) {
    CX acc = {0, 0};
    for (int j = 0; j < D; ++j) acc = FN(cadd)(acc, FN(cmul)(M[i * D + j], x[j]));
    y[i] = acc;
  }
}
static void FN(matvec_h)(int D, const CX* M, const CX* x, CX* y) { /* y = M^dag x */
  for (int i = 0; i < D; ++i) {
    CX acc = {0, 0};
    for (int j = 0; j < D; ++j) acc = FN(cadd)(acc, FN(cmul)(FN(cconj)(M[j * D + i]), x[j]));
    y[i] = acc;
  }
}

typedef struct {
  int D;
  const CX* R;      /* effective R */
  const float* f;   /* effective freqs (float32 values) */
  REAL A;
  REAL cterm;       /* -delta_t*sigma^2 as the reference forms it, model.py:312 (then /2.) */
  REAL delta_t;
} FN(model);

/* one _update_ancilla_psi (model.py:300-317): returns un-normalised psi' ; also u=R chi, chi */
static void FN(update)(const FN(model)* m, const CX* psi, REAL inc, const CX* p, CX* chi, CX* u, CX* v,
                       CX* delta, CX* psin) {
  const int D = m->D;
  const REAL s = inc / m->A;                                             /* :303 */
  for (int c = 0; c < D; ++c) chi[c] = FN(cmul)(psi[c], FN(cconj)(p[c])); /* :306 */
  FN(matvec)(D, m->R, chi, u);                                           /* :309 */
  FN(matvec_h)(D, m->R, u, v);                                           /* :310 */
  for (int c = 0; c < D; ++c) {
    CX d = FN(cscale)(FN(cscale)(v[c], m->cterm), (REAL)0.5);            /* :312  (-dt s^2 v)/2 */
    d = FN(cadd)(d, FN(cscale)(u[c], s));                                /* :313 */
    delta[c] = d;
    psin[c] = FN(cadd)(psi[c], FN(cmul)(p[c], d));                       /* :315-317 */
  }
}

/* _expectation (model.py:319-325) on psi at phases p; also returns chi' and w = R chi' */
static REAL FN(expectation)(const FN(model)* m, const CX* psi, const CX* p, CX* chi, CX* w) {
  const int D = m->D;
  for (int c = 0; c < D; ++c) chi[c] = FN(cmul)(psi[c], FN(cconj)(p[c]));
  FN(matvec)(D, m->R, chi, w);
  REAL acc = 0;
  for (int c = 0; c < D; ++c) acc += chi[c].re * w[c].re + chi[c].im * w[c].im; /* Re(conj(chi) w) */
  return 2 * acc;
}

/* _normalize_psi (model.py:327-334): returns the norm n = 1/rsqrt(max(sum |x|^2, eps)) */
static REAL FN(normalize)(int D, CX* x) {
  REAL ss = 0;
  for (int c = 0; c < D; ++c) {
#if IS_F32
    float a = hypotf(x[c].re, x[c].im);
#else
    double a = hypot(x[c].re, x[c].im);
#endif
    ss += a * a;
  }
  if (ss < (REAL)1e-12) ss = (REAL)1e-12;
#if IS_F32
  float inv = 1.0f / sqrtf(ss);
#else
  double inv = 1.0 / sqrt(ss);
#endif
  for (int c = 0; c < D; ++c) x[c] = FN(cscale)(x[c], inv);
  return 1 / inv;
}

static void FN(load_model)(FN(model)* m, int D, const float* R, const float* freqs, float A, float sigma,
                           double delta_t, CX* Rbuf) {
  for (int i = 0; i < D * D; ++i) { Rbuf[i].re = R[2 * i]; Rbuf[i].im = R[2 * i + 1]; }
  m->D = D; m->R = Rbuf; m->f = freqs; m->A = A;
  m->cterm = (REAL)(-delta_t * (double)sigma * (double)sigma);
  m->delta_t = (REAL)delta_t;
}

/* ---- per-clip loss (model.py:257-267, 276-282) -------------------------------------------- */
int FN(cmps_psi_loss)(int D, int B, int T, const float* R, const float* freqs, const float* psi0,
                      float A, float sigma, double delta_t, const float* x, double* loss) {
  CX* Rbuf = (CX*)malloc(sizeof(CX) * D * D);
  FN(model) m;
  FN(load_model)(&m, D, R, freqs, A, sigma, delta_t, Rbuf);
  float* tt = (float*)malloc(sizeof(float) * (T > 0 ? T : 1));
  ttable(T, delta_t, tt);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    CX* w8 = (CX*)malloc(sizeof(CX) * D * 8);
    CX *psi = w8, *p = w8 + D, *chi = w8 + 2 * D, *u = w8 + 3 * D, *v = w8 + 4 * D, *dl = w8 + 5 * D,
       *pn = w8 + 6 * D, *w = w8 + 7 * D;
    for (int c = 0; c < D; ++c) { psi[c].re = psi0[2 * c]; psi[c].im = psi0[2 * c + 1]; }
    REAL l = 0;
    const float* xb = x + (size_t)b * T;
    for (int k = 0; k + 1 < T; ++k) {
      const REAL inc = (REAL)xb[k + 1] - (REAL)xb[k];                    /* :263 */
      FN(phases)(D, m.f, tt[k], p);
      FN(update)(&m, psi, inc, p, chi, u, v, dl, pn);                    /* :278 */
      const REAL E = FN(expectation)(&m, pn, p, chi, w);                 /* :279, 293-294 */
#if IS_F32
      l += -logf(1.0f + (E * inc) / m.A);
#else
      l += -log(1.0 + (E * inc) / m.A);
#endif
      FN(normalize)(D, pn);                                              /* :280 */
      memcpy(psi, pn, sizeof(CX) * D);
    }
    loss[b] = (double)l;
    free(w8);
  }
  free(tt);
  free(Rbuf);
  return 0;
}

/* ---- loss + analytic adjoint, gradients wrt EFFECTIVE parameters ---------
 * L = sum_b w_b loss_b ;  gR[D*D*2], gpsi0[D*2] hold (dL/dRe, dL/dIm) pairs. */
int FN(cmps_psi_loss_grad)(int D, int B, int T, const float* R, const float* freqs, const float* psi0,
                           float A, float sigma, double delta_t, const float* x, const double* wts,
                           double* loss, double* gR, double* gf, double* gpsi0, double* gA) {
  CX* Rbuf = (CX*)malloc(sizeof(CX) * D * D);
  FN(model) m;
  FN(load_model)(&m, D, R, freqs, A, sigma, delta_t, Rbuf);
  float* tt = (float*)malloc(sizeof(float) * (T > 0 ? T : 1));
  ttable(T, delta_t, tt);
  const int nsteps = T > 0 ? T - 1 : 0;
  const size_t ng = (size_t)2 * D * D + D + 2 * D + 1;
  double* gall = (double*)calloc(ng * (size_t)B, sizeof(double));
  const REAL cprime = m.cterm * (REAL)0.5;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    double* g = gall + ng * (size_t)b;
    double *gRb = g, *gfb = g + 2 * D * D, *gpb = gfb + D, *gAb = gpb + 2 * D;
    CX* traj = (CX*)malloc(sizeof(CX) * (size_t)D * (nsteps + 1));       /* psi_k, normalised */
    CX* w12 = (CX*)malloc(sizeof(CX) * D * 16);
    CX *p = w12, *chi = w12 + D, *u = w12 + 2 * D, *v = w12 + 3 * D, *dl = w12 + 4 * D, *pn = w12 + 5 * D,
       *w = w12 + 6 * D, *chip = w12 + 7 * D, *lam = w12 + 8 * D, *gpp = w12 + 9 * D, *gchi = w12 + 10 * D,
       *gdel = w12 + 11 * D, *gu = w12 + 12 * D, *gv = w12 + 13 * D, *tmp = w12 + 14 * D, *gp = w12 + 15 * D;
    for (int c = 0; c < D; ++c) { traj[c].re = psi0[2 * c]; traj[c].im = psi0[2 * c + 1]; }
    const float* xb = x + (size_t)b * T;
    REAL l = 0;
    for (int k = 0; k < nsteps; ++k) {
      const REAL inc = (REAL)xb[k + 1] - (REAL)xb[k];
      FN(phases)(D, m.f, tt[k], p);
      FN(update)(&m, traj + (size_t)k * D, inc, p, chi, u, v, dl, pn);
      const REAL E = FN(expectation)(&m, pn, p, chip, w);
#if IS_F32
      l += -logf(1.0f + (E * inc) / m.A);
#else
      l += -log(1.0 + (E * inc) / m.A);
#endif
      FN(normalize)(D, pn);
      memcpy(traj + (size_t)(k + 1) * D, pn, sizeof(CX) * D);
    }
    loss[b] = (double)l;
    const REAL wb = (REAL)wts[b];
    for (int c = 0; c < D; ++c) { lam[c].re = 0; lam[c].im = 0; }
    for (int k = nsteps - 1; k >= 0; --k) {
      const CX* psi = traj + (size_t)k * D;
      const REAL inc = (REAL)xb[k + 1] - (REAL)xb[k];
      const REAL s = inc / m.A;
      FN(phases)(D, m.f, tt[k], p);
      FN(update)(&m, psi, inc, p, chi, u, v, dl, pn);           /* recompute psi', chi, u, v, delta */
      const REAL E = FN(expectation)(&m, pn, p, chip, w);       /* chi', w = R chi' */
      REAL n2 = 0;
      for (int c = 0; c < D; ++c) n2 += pn[c].re * pn[c].re + pn[c].im * pn[c].im;
#if IS_F32
      const REAL n = sqrtf(n2);
#else
      const REAL n = sqrt(n2);
#endif
      const REAL arg = 1 + (E * inc) / m.A;
      /* (a) through the normalisation: g_psi' = (lam - psi+ Re(psi+^dag lam)) / n */
      REAL dot = 0;
      for (int c = 0; c < D; ++c) {
        const CX pp = FN(cscale)(pn[c], 1 / n);
        dot += pp.re * lam[c].re + pp.im * lam[c].im;
      }
      for (int c = 0; c < D; ++c) {
        const CX pp = FN(cscale)(pn[c], 1 / n);
        gpp[c].re = (lam[c].re - pp.re * dot) / n;
        gpp[c].im = (lam[c].im - pp.im * dot) / n;
      }
      /* (b) loss term */
      const REAL gE = wb * (-(inc / m.A) / arg);
      *gAb += (double)(wb * (E * inc / (m.A * m.A)) / arg);
      FN(matvec_h)(D, m.R, chip, tmp);                          /* R^dag chi' */
      for (int c = 0; c < D; ++c) {
        gchi[c].re = 2 * gE * (w[c].re + tmp[c].re);
        gchi[c].im = 2 * gE * (w[c].im + tmp[c].im);
      }
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {                           /* gR += 2 gE chi' chi'^dag */
          const CX o = FN(cmul)(chip[i], FN(cconj)(chip[j]));
          gRb[2 * (i * D + j)] += (double)(2 * gE * o.re);
          gRb[2 * (i * D + j) + 1] += (double)(2 * gE * o.im);
        }
      /* (c) chi' = psi' conj(p) */
      for (int c = 0; c < D; ++c) {
        gpp[c] = FN(cadd)(gpp[c], FN(cmul)(p[c], gchi[c]));
        gp[c] = FN(cmul)(FN(cconj)(gchi[c]), pn[c]);
      }
      /* (d) psi' = psi + p delta */
      for (int c = 0; c < D; ++c) {
        gdel[c] = FN(cmul)(FN(cconj)(p[c]), gpp[c]);
        gp[c] = FN(cadd)(gp[c], FN(cmul)(gpp[c], FN(cconj)(dl[c])));
      }
      /* (e) delta = c' v + s u */
      REAL gs = 0;
      for (int c = 0; c < D; ++c) {
        gv[c] = FN(cscale)(gdel[c], cprime);
        gu[c] = FN(cscale)(gdel[c], s);
        gs += gdel[c].re * u[c].re + gdel[c].im * u[c].im;
      }
      *gAb += (double)(gs * (-inc / (m.A * m.A)));
      /* (f) v = R^dag u */
      FN(matvec)(D, m.R, gv, tmp);
      for (int c = 0; c < D; ++c) gu[c] = FN(cadd)(gu[c], tmp[c]);
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {                           /* gR += u gv^dag */
          const CX o = FN(cmul)(u[i], FN(cconj)(gv[j]));
          gRb[2 * (i * D + j)] += (double)o.re;
          gRb[2 * (i * D + j) + 1] += (double)o.im;
        }
      /* (g) u = R chi */
      FN(matvec_h)(D, m.R, gu, tmp);                            /* g_chi */
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {                           /* gR += gu chi^dag */
          const CX o = FN(cmul)(gu[i], FN(cconj)(chi[j]));
          gRb[2 * (i * D + j)] += (double)o.re;
          gRb[2 * (i * D + j) + 1] += (double)o.im;
        }
      /* (h) chi = psi conj(p) ; (d) g_psi = g_psi' */
      for (int c = 0; c < D; ++c) {
        lam[c] = FN(cadd)(gpp[c], FN(cmul)(p[c], tmp[c]));
        gp[c] = FN(cadd)(gp[c], FN(cmul)(FN(cconj)(tmp[c]), psi[c]));
      }
      /* (i) p = exp(i f t): gf += -t Im(conj(gp) p) */
      for (int c = 0; c < D; ++c) {
        const CX o = FN(cmul)(FN(cconj)(gp[c]), p[c]);
        gfb[c] += (double)(-(REAL)tt[k] * o.im);
      }
    }
    for (int c = 0; c < D; ++c) { gpb[2 * c] = lam[c].re; gpb[2 * c + 1] = lam[c].im; }
    free(w12);
    free(traj);
  }
  for (size_t e = 0; e < ng; ++e) {
    double s = 0;
    for (int b = 0; b < B; ++b) s += gall[ng * (size_t)b + e];
    if (e < (size_t)2 * D * D) gR[e] = s;
    else if (e < (size_t)2 * D * D + D) gf[e - 2 * D * D] = s;
    else if (e < (size_t)2 * D * D + 3 * D) gpsi0[e - 2 * D * D - D] = s;
    else *gA = s;
  }
  free(gall);
  free(tt);
  free(Rbuf);
  return 0;
}

/* ---- sampler (model.py:242-251, 284-291); noise [L][n] -> out [n][L] -------------------------- */
int FN(cmps_psi_sample)(int D, int n, int L, const float* R, const float* freqs, const float* psi0,
                        float A, float sigma, double delta_t, const float* noise, double* out) {
  CX* Rbuf = (CX*)malloc(sizeof(CX) * D * D);
  FN(model) m;
  FN(load_model)(&m, D, R, freqs, A, sigma, delta_t, Rbuf);
  float* tt = (float*)malloc(sizeof(float) * (L + 1));
  ttable(L + 1, delta_t, tt);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < n; ++b) {
    CX* w8 = (CX*)malloc(sizeof(CX) * D * 8);
    CX *psi = w8, *p = w8 + D, *chi = w8 + 2 * D, *u = w8 + 3 * D, *v = w8 + 4 * D, *dl = w8 + 5 * D,
       *pn = w8 + 6 * D, *w = w8 + 7 * D;
    for (int c = 0; c < D; ++c) { psi[c].re = psi0[2 * c]; psi[c].im = psi0[2 * c + 1]; }
    REAL X = 0;
    for (int k = 0; k < L; ++k) {
      FN(phases)(D, m.f, tt[k], p);
      const REAL E = FN(expectation)(&m, psi, p, chi, w);
      const REAL inc = E * m.delta_t + (REAL)noise[(size_t)k * n + b];   /* :286 */
      X += inc;                                                          /* :287 */
      FN(update)(&m, psi, inc, p, chi, u, v, dl, pn);                    /* :288 */
      FN(normalize)(D, pn);                                              /* :289 */
      memcpy(psi, pn, sizeof(CX) * D);
      out[(size_t)b * L + k] = (double)(m.A * X);                        /* :251 */
    }
    free(w8);
  }
  free(tt);
  free(Rbuf);
  return 0;
}

#undef CX
#undef FN
#undef CAT
#undef CAT_
