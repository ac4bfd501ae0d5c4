// Gradient tiles of the adjoint sweep on the 5th-generation tensor cores (D = 33..64).
//
// The three rank-1 sums of the backward (DESIGN.md "Adjoint"),
//     G_N = sum_k mu_k x_k^dag,   G_R = sum_k s_k mu_k x_k^dag,   G_E = sum_k alpha_k x'_k x'_k^dag,
// do not feed the adjoint recursion: inside the sequential kernel they were 96 of ~150 FFMA per thread and
// step plus two thirds of its shared-memory loads (the D = 64 backward sat at 80 % of the LSU pipe,
// profiles/r1_ncu_uni.md).  Over the time axis they are GEMMs with K = T: the sequential kernel now only
// runs the chain and stores mu_k (in place of the S x'_k it has just consumed), and THIS kernel contracts
// the stored trajectories, fully parallel over (clip, time range):
//     kind 0 (D <= 64):  [G_N | G_R]  = M [X ; X diag(s)]^T      (M = mu,      UMMA 128 x 256, K = steps)
//     kind 1          :   G_E         = (X' diag(alpha)) X'^T    (X' = x',     UMMA 128 x 2D)
//     kind 2, 3 (D = 128): G_N = M X^T and G_R = M (X diag(s))^T separately (UMMA 128 x 256 each)
// A CTA owns 64 complex rows of the result (one UMMA M = 128 tile): at D = 128 blockIdx.y selects the half.
// Complex D x K operands are fed as REAL 2D x K matrices (row 2r = Re, 2r+1 = Im of component r); the
// real 2D x 2D product holds all four Re/Im pairings and the epilogue folds them,
//     G[r][r'] = (C[2r][2r'] + C[2r+1][2r'+1]) + i (C[2r+1][2r'] - C[2r][2r'+1]).
// kind::tf32 with the 3-pass hi/lo split (A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in tensor
// memory): per-product error 2^-21 instead of 2^-10, so the tiles keep float32-level accuracy.
// Operands are K-major, 128-byte swizzled, staged by 15 producer warps straight from the global
// trajectories (x'_k = conj(q_k) x_{k+1} / c_k, s_k, alpha_k are formed on the way: nothing else is stored),
// two stages of 32 steps; one thread issues the MMAs; tcgen05.commit -> mbarrier frees a stage.
#pragma once
#include "amps_common.cuh"
#include "amps_scan_tc.cuh"

namespace amps {

constexpr int TL_KS = 32;            // steps per stage = floats per 128-byte swizzle row
constexpr int TL_STAGES = 2;
constexpr int TL_THREADS = 512;      // producer threads (warps 0..15; warps 0..3 also run the epilogue)
constexpr int TL_BLOCK = TL_THREADS + 32;   // + warp 16: issues the MMAs, nothing else (a tcgen05.mma can hold its
                                            // thread while the pipe's queue is full: it must not gate the staging)
constexpr int TL_ROWB = 128;         // bytes per operand row

template <int DP, int TYPE>
struct alignas(1024) TilesSmem {
  static constexpr int MA = 128;                          // A rows (real form of 64 complex rows)
  static constexpr int NB = TYPE == 0 ? 4 * DP : 2 * DP;  // B rows
  uint8_t a_hi[TL_STAGES][MA * TL_ROWB];
  uint8_t a_lo[TL_STAGES][MA * TL_ROWB];
  uint8_t b_hi[TL_STAGES][NB * TL_ROWB];
  uint8_t b_lo[TL_STAGES][NB * TL_ROWB];
  unsigned long long full_bar[TL_STAGES];
  unsigned long long empty_bar[TL_STAGES];
  unsigned long long done_bar;
  uint32_t tmem_base;
};

// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void tl_mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tl_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(tc_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tl_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc_smem_u32(bar)) : "memory");
}
// byte offset of (row n, 16-byte chunk c of the 128-byte row) in a K-major SWIZZLE_128B tile
__device__ __forceinline__ int tl_off(int n, int c) { return n * TL_ROWB + ((c ^ (n & 7)) << 4); }

struct TilesArgs {
  const float2* mu;       // [B][T][DP]  adjoint of x'_k in row k            (type 0)
  const float2* traj;     // [B][T][DP]  x_k (scaled), row k
  const float2* qtab;     // [nsteps][DP]                                     (type 1)
  const float* scales;    // [B][nchunks] c_k of each rescale chunk           (type 1)
  const float2* ev;       // [B][T] (alpha_k, 1/c_k), left by the chain-only adjoint sweep   (type 1)
  const float* x;         // waveform, clip stride xstride
  const float* w;         // [B] clip weights                                  (type 1)
  float2* G;              // [B * nsplit][3][DP][DP] partial tiles: [0] = G_R, [1] = G_N, [2] = G_E
  int T, xstride, nchunks, chunk_len;
  int nsplit, steps_per_split;   // steps_per_split: multiple of TL_KS
  int accumulate;                // add to the values already in G (time windows of the checkpointed backward)
  AVal A;
};

// grid = (B * nsplit, DP / 64), block = 544
template <int DP, int TYPE>
__global__ void __launch_bounds__(TL_BLOCK, 1) psi_tiles_tc_kernel(TilesArgs g) {
  using Sm = TilesSmem<DP, TYPE>;
  constexpr int MA = Sm::MA, NB = Sm::NB;
  static_assert(NB <= 256 && (DP == 64 || DP == 128), "UMMA N <= 256");
  constexpr int RA = 64;                       // complex rows of the result owned by this CTA
  const int ra0 = blockIdx.y * RA;             // first of them (D = 128: two halves)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem_al = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  Sm& sm = *reinterpret_cast<Sm*>(smem_al);
  const float A = a_get(g.A);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / g.nsplit, sp = blockIdx.x % g.nsplit;
  const int nsteps = g.T - 1;
  const int k_begin = sp * g.steps_per_split;
  const int nloc = max(0, min(g.steps_per_split, nsteps - k_begin));
  const int nblk = (nloc + TL_KS - 1) / TL_KS;
  const size_t rowbase = (size_t)b * g.T;

  if (tid == 0) {
    for (int s = 0; s < TL_STAGES; ++s) {
      mbar_init(&sm.full_bar[s], TL_THREADS);
      mbar_init(&sm.empty_bar[s], 1);
    }
    mbar_init(&sm.done_bar, 1);
    mbar_fence_init_cluster();
  }
  constexpr uint32_t NCOLS = NB;   // fp32 accumulator: one tensor-memory column per B row
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc_smem_u32(&sm.tmem_base)),
                 "r"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = NB, M = 128
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(MA >> 4) << 24);

  // ---- producers (all 512 threads).  A block of 32 steps is staged in two halves: fetch(j) issues every
  // global load of block j into registers, stage(j, s) turns them into operand tiles.  fetch(j+1) is issued
  // right after stage(j), so the loads are in flight while the stage's MMAs run (a plain load-then-use pass
  // left the kernel bound by one exposed HBM latency per block).
  constexpr int NBT = DP * (TL_KS / 4) / TL_THREADS;   // B tasks per thread: (component, group of 4 steps)
  static_assert(NBT * TL_THREADS == DP * (TL_KS / 4) && RA * (TL_KS / 4) == TL_THREADS, "static task map");
  // Task map: a warp covers 16 components x 2 step groups -- lanes 4..7 of every eight take the SAME four
  // components as lanes 0..3 but the neighbouring step group (kg ^ 1).  The Re rows (2r: rows 0,2,4,6 mod 8)
  // of a quarter warp then land on eight different 16-byte bank groups of the swizzled tile (chunk kg or
  // kg ^ 1, xor the row), and so do the Im rows: both operand stores are conflict free without any select.
  const int l_r = (lane & 3) + 4 * (lane >> 3), l_kg = (lane >> 2) & 1;
  auto task_of = [&](int wv, int ncomp, int& r, int& kg) {   // wv: warp index over the task set
    const int per = ncomp / 16;
    r = 16 * (wv % per) + l_r;
    kg = 2 * (wv / per) + l_kg;
  };
  struct Pre {              // one block's global data, per thread
    float2 pb[NBT][4];      // kind 0, 2, 3: x_k ; kind 1: x_{k+1}
    float2 pq[TYPE == 1 ? NBT : 1][4];   // kind 1: q_k
    float2 pa[4];           // kind 0, 2, 3: mu_k of this CTA's component
    float xw[TYPE == 1 ? 1 : NBT][5];    // kind 0, 3: waveform samples of the task's four steps (s_k)
    float2 ev[TYPE == 1 ? NBT : 1][4];   // kind 1: (alpha_k, 1 / c_k)
  };
  const float* xb = g.x + (size_t)b * g.xstride;
  auto fetch = [&](Pre& P, int j) {
    const int k0 = k_begin + j * TL_KS;
    const int len = min(TL_KS, k_begin + nloc - k0);
#pragma unroll
    for (int n = 0; n < NBT; ++n) {
      int r, kg;
      task_of(warp + 16 * n, DP, r, kg);
      if (TYPE == 0 || TYPE == 3) {
#pragma unroll
        for (int e = 0; e < 5; ++e) P.xw[n][e] = (4 * kg + e <= len) ? xb[k0 + 4 * kg + e] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kk = 4 * kg + e;
        P.pb[n][e] = make_float2(0.f, 0.f);
        if (TYPE == 1) {
          P.pq[n][e] = make_float2(0.f, 0.f);
          P.ev[n][e] = make_float2(0.f, 0.f);
        }
        if (kk < len) {
          const int k = k0 + kk;
          if (TYPE != 1) {
            P.pb[n][e] = g.traj[(rowbase + k) * DP + r];
          } else {
            P.pq[n][e] = g.qtab[(size_t)k * DP + r];
            P.pb[n][e] = g.traj[(rowbase + k + 1) * DP + r];
            P.ev[n][e] = g.ev[rowbase + k];
          }
        }
      }
    }
    if (TYPE != 1) {
      int rl, kg;
      task_of(warp, RA, rl, kg);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kk = 4 * kg + e;
        P.pa[e] = (kk < len) ? g.mu[(rowbase + k0 + kk) * DP + ra0 + rl] : make_float2(0.f, 0.f);
      }
    }
  };
  auto put2 = [&](uint8_t* hi, uint8_t* lo, int r, const float (&re)[4], const float (&im)[4], int kg) {
    const float4 rh = make_float4(tc_trunc_tf32(re[0]), tc_trunc_tf32(re[1]), tc_trunc_tf32(re[2]), tc_trunc_tf32(re[3]));
    const float4 ih = make_float4(tc_trunc_tf32(im[0]), tc_trunc_tf32(im[1]), tc_trunc_tf32(im[2]), tc_trunc_tf32(im[3]));
    const int o0 = tl_off(2 * r, kg), o1 = tl_off(2 * r + 1, kg);
    *reinterpret_cast<float4*>(hi + o0) = rh;
    *reinterpret_cast<float4*>(lo + o0) = make_float4(re[0] - rh.x, re[1] - rh.y, re[2] - rh.z, re[3] - rh.w);
    *reinterpret_cast<float4*>(hi + o1) = ih;
    *reinterpret_cast<float4*>(lo + o1) = make_float4(im[0] - ih.x, im[1] - ih.y, im[2] - ih.z, im[3] - ih.w);
  };
  const float rcpA = 1.0f / A;
  auto stage = [&](const Pre& P, int j, int s) {
    // B operand (all DP components); kind 1 also fills the A rows of this CTA's components from the same x'
#pragma unroll
    for (int n = 0; n < NBT; ++n) {
      int r, kg;
      task_of(warp + 16 * n, DP, r, kg);
      float sk[4];          // kind 0, 3: s_k ; kind 1: alpha_k
      float bre[4], bim[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (TYPE != 1) {
          if (TYPE != 2) sk[e] = (P.xw[n][e + 1] - P.xw[n][e]) * rcpA;   // s_k = inc_k / A
          bre[e] = P.pb[n][e].x;
          bim[e] = P.pb[n][e].y;
        } else {
          sk[e] = P.ev[n][e].x;                                            // alpha_k
          const float2 xp = cmul_ca(P.pq[n][e], P.pb[n][e]);               // x'_k = conj(q_k) x_{k+1} / c_k
          bre[e] = xp.x * P.ev[n][e].y;
          bim[e] = xp.y * P.ev[n][e].y;
        }
      }
      if (TYPE == 0 || TYPE == 3) {   // s_k x_k
        float sre[4], sim[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sre[e] = sk[e] * bre[e];
          sim[e] = sk[e] * bim[e];
        }
        // kind 0: stacked under x (rows 2 DP ..) ; kind 3: the only B rows
        put2(sm.b_hi[s] + (TYPE == 0 ? 2 * DP * TL_ROWB : 0), sm.b_lo[s] + (TYPE == 0 ? 2 * DP * TL_ROWB : 0), r, sre, sim, kg);
      }
      if (TYPE != 3) put2(sm.b_hi[s], sm.b_lo[s], r, bre, bim, kg);
      if (TYPE == 1 && r >= ra0 && r < ra0 + RA) {   // A = alpha_k x'_k  (warp-uniform: 16 components per warp)
        float are[4], aim[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          are[e] = sk[e] * bre[e];
          aim[e] = sk[e] * bim[e];
        }
        put2(sm.a_hi[s], sm.a_lo[s], r - ra0, are, aim, kg);
      }
    }
    if (TYPE != 1) {   // A = mu_k, this CTA's 64 components
      int rl, kg;
      task_of(warp, RA, rl, kg);
      float are[4], aim[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        are[e] = P.pa[e].x;
        aim[e] = P.pa[e].y;
      }
      put2(sm.a_hi[s], sm.a_lo[s], rl, are, aim, kg);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the MMA
    tl_mbar_arrive(&sm.full_bar[s]);
  };

  // MMA thread: MMAs of block jm out of stage jm % 2
  auto issue_mma = [&](int jm) {
    const int s = jm % TL_STAGES;
    mbar_wait_cta(&sm.full_bar[s], (jm / TL_STAGES) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      const uint64_t da0 = tc_make_desc(tc_smem_u32(pass == 1 ? sm.a_lo[s] : sm.a_hi[s]));
      const uint64_t db0 = tc_make_desc(tc_smem_u32(pass == 2 ? sm.b_lo[s] : sm.b_hi[s]));
#pragma unroll
      for (int ks = 0; ks < TL_KS / 8; ++ks)      // UMMA_K = 8 tf32 = 32 bytes along the swizzled row
        tl_mma_ss(tmem, da0 + 2 * ks, db0 + 2 * ks, idesc, (jm > 0 || pass > 0 || ks > 0) ? 1u : 0u);
    }
    tl_commit(&sm.empty_bar[s]);
    if (jm == nblk - 1) tl_commit(&sm.done_bar);
  };
  // ---- main loop: warp 16 issues, warps 0..15 stage (register sets ping-pong) -------------------------
  if (warp == TL_THREADS / 32) {
    if (lane == 0)
      for (int jm = 0; jm < nblk; ++jm) issue_mma(jm);
  } else {
    // fetch block j+1 into `nxt` -- its loads fly while block j is staged out of `cur`.  (Where two register
    // sets do not fit, block j+1 is fetched into the same set right after block j has been staged: the other
    // three warps of the scheduler cover the load latency.)
    constexpr bool TWO_SETS = !(DP == 128 && TYPE == 1);
    auto iterate = [&](Pre& cur, Pre& nxt, int j) {
      if (TWO_SETS && j + 1 < nblk) fetch(nxt, j + 1);
      const int s = j % TL_STAGES;
      if (j >= TL_STAGES) mbar_wait_cta(&sm.empty_bar[s], ((j / TL_STAGES) - 1) & 1);   // MMAs of block j-2 done
      stage(cur, j, s);
      if (!TWO_SETS && j + 1 < nblk) fetch(cur, j + 1);
    };
    Pre P0;
    if (nblk > 0) fetch(P0, 0);
    if (TWO_SETS) {
      Pre P1;
      for (int j = 0; j < nblk; j += 2) {
        iterate(P0, P1, j);
        if (j + 1 < nblk) iterate(P1, P0, j + 1);
      }
    } else {
      for (int j = 0; j < nblk; ++j) iterate(P0, P0, j);
    }
  }

  // ---- epilogue: fold the four real pairings, write this CTA's partial tiles -----------------------
  float2* Gb = g.G + (size_t)blockIdx.x * 3 * DP * DP;
  if (warp < 4) {
    if (nblk > 0) {
      mbar_wait_cta(&sm.done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
    const int m = warp * 32 + lane;          // TMEM lane = real row 2 r_local + c of A
    const int r = ra0 + (m >> 1), c = m & 1;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < (int)NCOLS; c0 += 32) {
      float v[32];
      if (nblk > 0) {
        tc_ld32(lane_base + c0, v);
      } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
      }
      // columns c0 + 2 j', c0 + 2 j' + 1 = (Re, Im) row pair of B component rp
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const float p = __shfl_xor_sync(0xffffffffu, v[2 * jj + 1], 1);
        const float val = (c == 0) ? v[2 * jj] + p : v[2 * jj] - p;   // even lane: Re, odd lane: Im
        const int ncol = (c0 >> 1) + jj;                                // complex column index within [0, NB/2)
        int which, rp;
        if (TYPE == 0) {
          which = ncol < DP ? 1 : 0;   // B rows [0, 2DP) = x -> G_N ; [2DP, 4DP) = s x -> G_R
          rp = ncol % DP;
        } else {
          which = TYPE == 1 ? 2 : (TYPE == 2 ? 1 : 0);
          rp = ncol;
        }
        float* dst = reinterpret_cast<float*>(Gb + (size_t)which * DP * DP + (size_t)r * DP + rp) + c;
        *dst = g.accumulate ? *dst + val : val;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(NCOLS));
}

}  // namespace amps
