// Parallel-in-time scan for small batches (K4), forward and adjoint: the clip is cut into `nvc` time chunks;
//   1. psi_compose_tc_kernel  -- one CTA per (clip, chunk) composes the chunk's step operators
//        C_j = W_{k0+m-1} ... W_{k0},   W_k = diag(q_k) (I + E_k),   E_k = c' R^dag R + s_k R
//      on the 5th-generation tensor cores: complex D x D operators as real 2D x 2D matrices
//      (D = 64 -> UMMA M = N = K = 128, kind::tf32).  The running product lives in TENSOR MEMORY
//      and is fed back as the A operand (TS mode); E_k is formed in shared memory (K-major,
//      128-byte swizzle) as a tf32 hi/lo pair, and three MMA passes (P_hi E_hi + P_lo E_hi + P_hi E_lo)
//      accumulate P^T E^T ON TOP of an exact fp32 copy of P^T, so only the small correction E P
//      (|E| ~ 3e-2) goes through tf32 products: per-step error ~1e-8 |P|.  The diagonal phase
//      rotation q_k is applied in fp32 by the epilogue warps (tcgen05.ld / tcgen05.st).
//   2. psi_scan_boundary_kernel -- sequential over the nvc chunk operators: chunk start states.
//   3. the sequential forward kernel replays every chunk from its start state as an independent
//      "virtual clip" (loss terms, optional trajectory).
//   backward (amps_psi_loss_bwd_scan): 4. virtual-clip adjoint with a zero end condition (chain only) ->
//      d_j;  5. psi_scan_boundary_bwd_kernel: Lam_j = d_j + C_j^dag Lam_{j+1} / |C_j y_j|, sequential over
//      the stored chunk operators;  6. virtual-clip adjoint from the true end adjoints (gradient tiles).
// Costs 8 D^3 flops per step instead of 24 D^2 (x21 at D = 64) but turns ONE latency-bound chain of
// T steps into 148 chains of T/148 steps: worthwhile only when the batch cannot fill the GPU.
#pragma once
#include "amps_common.cuh"

namespace amps {

constexpr int TC_D = 64;               // complex bond dimension handled here (smaller D is zero-padded)
constexpr int TC_N = 2 * TC_D;         // real-form dimension = UMMA M = N = K
constexpr int TC_KB = 32;              // floats per 128-byte swizzle row
constexpr int TC_NKB = TC_N / TC_KB;   // K blocks per operand
constexpr int TC_TILE = TC_N * 128;    // bytes of one K block (128 rows x 128 B)

struct alignas(1024) ScanTcSmem {
  uint8_t bhi[TC_NKB * TC_TILE];   // E_k truncated to tf32      (real form, K-major, SWIZZLE_128B)
  uint8_t blo[TC_NKB * TC_TILE];   // E_k - trunc(E_k)
  float2 cM[TC_D][TC_D];           // c' R^dag R  ( = N - I )
  float2 Rm[TC_D][TC_D];
  float2 qv[2][TC_D];              // q_k, by step parity
  unsigned long long mbar;         // all MMAs of a step are complete
  unsigned long long mbar_lo;      // the E_lo pass of a step is complete (blo may be rewritten)
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout)
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address            [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset      [16,30) (unused: swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset       [32,46): 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                          uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ float tc_trunc_tf32(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

// TMEM column map (512 columns allocated): X = running product P^T (A operand), Z = its tf32 residual,
// Dacc = exact fp32 copy of P^T that the MMAs accumulate P^T E^T onto.
constexpr uint32_t TC_COL_X = 0, TC_COL_Z = 128, TC_COL_D = 256;

// grid = B * nvc CTAs, block = 512 threads: warps 0-7 run the per-step epilogue (warp w: TMEM lanes 32 (w & 3) ..,
// column-pair half w >> 2), warps 8-15 form E_k in shared memory; thread 0 issues the MMAs.
constexpr int TC_THREADS = 512;
constexpr int TC_EPI = 256;      // threads of the per-step epilogue (warps 0-7); the rest form E_k
// HALF (bond dimension <= 32): E_k is zero outside rows/columns [0,32) of each real-form quadrant, i.e.
// K blocks 1 and 3 of the B operand never change from zero -- they are neither re-formed nor multiplied.
template <bool HALF>
__global__ void __launch_bounds__(TC_THREADS)
    psi_compose_tc_kernel(const float2* __restrict__ matN, const float2* __restrict__ matR,
                          const float2* __restrict__ qtab, const float* __restrict__ x, int T, AVal A_,
                          int nvc, int m_steps, float* __restrict__ opsT) {
  const float A = a_get(A_);
  constexpr int Dq = HALF ? 32 : TC_D;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // SWIZZLE_128B operands need a 1024-byte aligned base (the launch reserves 1 KB of slack)
  unsigned char* smem_al = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  ScanTcSmem& sm = *reinterpret_cast<ScanTcSmem*>(smem_al);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int vc = blockIdx.x, clip = vc / nvc, jc = vc % nvc;
  const int nsteps = T - 1;
  const int k_begin = jc * m_steps;
  const int nloc = max(0, min(m_steps, nsteps - k_begin));
  const float* xb = x + (size_t)clip * T + k_begin;
  const float2* qb = qtab + (size_t)k_begin * TC_D;

  for (int idx = tid; idx < TC_D * TC_D; idx += TC_THREADS) {
    const int a = idx / TC_D, c = idx % TC_D;
    float2 n = matN[idx];
    if (a == c) n.x -= 1.0f;          // c' R^dag R = N - I (exact for N_aa ~ 1)
    sm.cM[a][c] = n;
    sm.Rm[a][c] = matR[idx];
  }
  if (Dq < TC_D) {
    for (int idx = tid; idx < TC_NKB * TC_TILE / 16; idx += TC_THREADS) {
      reinterpret_cast<float4*>(sm.bhi)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(sm.blo)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (tid == 0) {
    mbar_init(&sm.mbar, 1);
    mbar_init(&sm.mbar_lo, 1);
    mbar_fence_init_cluster();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(tc_smem_u32(&sm.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);   // (epilogue warps)

  if (tid < 128) {   // P^T = I : X = I, Dacc = I, Z = 0
    float v[32], z[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = 0.f;
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (c0 + j == tid) ? 1.0f : 0.0f;
      tc_st32(lane_base + TC_COL_X + c0, v);
      tc_st32(lane_base + TC_COL_D + c0, v);
      tc_st32(lane_base + TC_COL_Z + c0, z);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }

  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_N >> 4) << 24);

  // ---- producers: E_k = c' R^dag R + s_k R in real form [[Er,-Ei],[Ei,Er]], tf32 hi / lo, swizzled --------------
  // One thread per (row a, 4 consecutive columns b): the complex element is computed once and written to its
  // four real-form positions, hi and lo.  The MMAs of a step run the E_lo pass FIRST, so blo is free after a third
  // of the step's tensor work: E_{k+1} is computed and its lo half written while the rest of step k's MMAs run;
  // only the hi half (kept in registers meanwhile) is written between the steps.  (Rewriting the whole 128 KB
  // operand between two steps was ~2 k of the 7.9 k cycles per step.)
  constexpr int NPROD = TC_THREADS - TC_EPI;
  constexpr int cq = Dq / 4;
  constexpr int NITEM = (Dq * cq + NPROD - 1) / NPROD;
  const int pt = tid - TC_EPI;
  float4 keep_er[NITEM], keep_ei[NITEM];     // hi parts of the step being prepared
  auto boff = [](int n, int cch) { return (cch / 8) * TC_TILE + n * 128 + (((cch % 8) ^ (n & 7)) * 16); };
  auto produce_lo = [&](int k) {             // E_k: lo half -> blo, hi half -> registers
    const float s = (xb[k + 1] - xb[k]) / A;                           // model.py:263, 303
#pragma unroll
    for (int it = 0; it < NITEM; ++it) {
      const int idx = pt + it * NPROD;
      if (idx < Dq * cq) {
        const int a = idx / cq, c = idx % cq;                           // c: 16-byte chunk within [0, D)
        float er[4], ei[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 m = sm.cM[a][4 * c + e], r = sm.Rm[a][4 * c + e];
          er[e] = fmaf(s, r.x, m.x);
          ei[e] = fmaf(s, r.y, m.y);
        }
        const float4 erh = make_float4(tc_trunc_tf32(er[0]), tc_trunc_tf32(er[1]), tc_trunc_tf32(er[2]), tc_trunc_tf32(er[3]));
        const float4 eih = make_float4(tc_trunc_tf32(ei[0]), tc_trunc_tf32(ei[1]), tc_trunc_tf32(ei[2]), tc_trunc_tf32(ei[3]));
        const float4 erl = make_float4(er[0] - erh.x, er[1] - erh.y, er[2] - erh.z, er[3] - erh.w);
        const float4 eil = make_float4(ei[0] - eih.x, ei[1] - eih.y, ei[2] - eih.z, ei[3] - eih.w);
        keep_er[it] = erh;
        keep_ei[it] = eih;
        const int cl = c, cr = c + TC_D / 4, rt = a, rb = a + TC_D;   // chunk within a 128-column row: left c, right c + 16
        *reinterpret_cast<float4*>(sm.blo + boff(rt, cl)) = erl;                                        // top-left      Er
        *reinterpret_cast<float4*>(sm.blo + boff(rt, cr)) = make_float4(-eil.x, -eil.y, -eil.z, -eil.w);  // top-right    -Ei
        *reinterpret_cast<float4*>(sm.blo + boff(rb, cl)) = eil;                                        // bottom-left   Ei
        *reinterpret_cast<float4*>(sm.blo + boff(rb, cr)) = erl;                                        // bottom-right  Er
      }
    }
  };
  auto produce_hi = [&]() {
#pragma unroll
    for (int it = 0; it < NITEM; ++it) {
      const int idx = pt + it * NPROD;
      if (idx < Dq * cq) {
        const int a = idx / cq, c = idx % cq;
        const float4 erh = keep_er[it], eih = keep_ei[it];
        const int cl = c, cr = c + TC_D / 4, rt = a, rb = a + TC_D;
        *reinterpret_cast<float4*>(sm.bhi + boff(rt, cl)) = erh;
        *reinterpret_cast<float4*>(sm.bhi + boff(rt, cr)) = make_float4(-eih.x, -eih.y, -eih.z, -eih.w);
        *reinterpret_cast<float4*>(sm.bhi + boff(rb, cl)) = eih;
        *reinterpret_cast<float4*>(sm.bhi + boff(rb, cr)) = erh;
      }
    }
  };
  if (tid >= TC_EPI && nloc > 0) produce_lo(0);

  for (int kk = 0; kk <= nloc; ++kk) {
    if (kk > 0) mbar_wait_cta(&sm.mbar, (kk - 1) & 1); // the MMAs of step kk-1 are complete (tcgen05.commit)
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid >= TC_EPI) {
      if (kk < nloc) {
        if (pt < TC_D) sm.qv[kk & 1][pt] = qb[(size_t)kk * TC_D + pt];
        produce_hi();                                   // E_kk, hi half (lo half is in place already)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      }
    } else if (kk > 0) {
      // ---- epilogue of step kk-1: Dacc = (P + E P)^T ; rotate column pairs (a, a+D) by q_a ----
      // (eight warps: warp w owns TMEM lanes 32 (w & 3) .. and the column-pair half h = w >> 2)
      const float2* q = sm.qv[(kk - 1) & 1];
      const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      {
        const int h = warp >> 2;
        float ya[32], yb[32];
        tc_ld32(lane_base + TC_COL_D + 32 * h, ya);
        tc_ld32(lane_base + TC_COL_D + TC_D + 32 * h, yb);
        float lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 qa = q[32 * h + j];
          const float va = qa.x * ya[j] - qa.y * yb[j];
          const float vb = fmaf(qa.y, ya[j], qa.x * yb[j]);
          ya[j] = va;
          yb[j] = vb;
        }
        tc_st32(lane_base + TC_COL_X + 32 * h, ya);
        tc_st32(lane_base + TC_COL_D + 32 * h, ya);
#pragma unroll
        for (int j = 0; j < 32; ++j) lo[j] = ya[j] - tc_trunc_tf32(ya[j]);
        tc_st32(lane_base + TC_COL_Z + 32 * h, lo);
        tc_st32(lane_base + TC_COL_X + TC_D + 32 * h, yb);
        tc_st32(lane_base + TC_COL_D + TC_D + 32 * h, yb);
#pragma unroll
        for (int j = 0; j < 32; ++j) lo[j] = yb[j] - tc_trunc_tf32(yb[j]);
        tc_st32(lane_base + TC_COL_Z + TC_D + 32 * h, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0 && kk < nloc) {
      // Dacc += P_hi^T E_lo^T (first: frees blo) + P_hi^T E_hi^T + P_lo^T E_hi^T   (A from tensor memory, B from smem)
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t acol = tmem + ((pass == 2) ? TC_COL_Z : TC_COL_X);
        const uint8_t* bsm = (pass == 0) ? sm.blo : sm.bhi;
#pragma unroll 1
        for (int kb = 0; kb < TC_NKB; ++kb) {
          if (Dq < TC_D && (kb & 1)) continue;          // all-zero K block of E_k^T
          const uint64_t db0 = tc_make_desc(tc_smem_u32(bsm + kb * TC_TILE));
#pragma unroll
          for (int ks = 0; ks < TC_KB / 8; ++ks)      // UMMA_K = 8 tf32: 8 TMEM columns / 32 smem bytes
            tc_mma_ts(tmem + TC_COL_D, acol + kb * TC_KB + ks * 8, db0 + 2 * ks, idesc, 1u);
        }
        if (pass == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                           tc_smem_u32(&sm.mbar_lo))
                       : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                       tc_smem_u32(&sm.mbar))
                   : "memory");
    }
    if (tid >= TC_EPI && kk + 1 < nloc) {
      // E_{kk+1}: computed now, its lo half written as soon as step kk's E_lo pass has been read
      mbar_wait_cta(&sm.mbar_lo, kk & 1);
      produce_lo(kk + 1);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
  }

  // chunk operator out: opsT[vc][m][n] = P^T[m][n]
  if (tid < 128) {
    float* dst = opsT + ((size_t)vc * TC_N + tid) * TC_N;
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
      float v[32];
      tc_ld32(lane_base + TC_COL_X + c0, v);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

// Sequential pass over the chunk operators of one clip: normalised chunk start states.
// grid = B, block = 256; the next chunk operator (64 KB) streams into shared memory with cp.async
// while the current mat-vec runs.
__global__ void __launch_bounds__(256)
    psi_scan_boundary_kernel(const float* __restrict__ opsT, const float2* __restrict__ psi0p, int nvc,
                             float2* __restrict__ ystart, float* __restrict__ rnv) {
  extern __shared__ __align__(16) unsigned char bsm_raw[];
  float* pt = reinterpret_cast<float*>(bsm_raw);          // [2][128][128]
  float* y = pt + 2 * TC_N * TC_N;                        // [128]
  float* part = y + TC_N;                                 // [2][128]
  float* red = part + 2 * TC_N;                           // [8]
  const int t = threadIdx.x, b = blockIdx.x;
  const int n = t & (TC_N - 1), half = t >> 7;            // column n, rows [64*half, 64*half+64)
  auto issue = [&](int j) {
    const float* src = opsT + ((size_t)b * nvc + j) * TC_N * TC_N;
    float* dst = pt + (j & 1) * TC_N * TC_N;
    for (int idx = t; idx < TC_N * TC_N / 4; idx += 256) cp_async16(dst + 4 * idx, src + 4 * idx);
  };
  if (t < TC_D) {
    const float2 p = psi0p[t];
    y[t] = p.x;
    y[t + TC_D] = p.y;
  }
  if (nvc > 0) issue(0);
  cp_async_commit();
  __syncthreads();
  for (int j = 0; j < nvc; ++j) {
    if (j + 1 < nvc) issue(j + 1);
    cp_async_commit();
    // normalise and emit the start state of chunk j
    float n2 = (t < TC_N) ? y[t] * y[t] : 0.f;
    n2 = warp_sum_f(n2);
    if ((t & 31) == 0) red[t >> 5] = n2;
    __syncthreads();
    const float rn = rsqrtf(fmaxf(red[0] + red[1] + red[2] + red[3], 1e-12f));
    if (t < TC_N) y[t] *= rn;
    cp_async_wait<1>();
    __syncthreads();
    if (t < TC_D) ystart[((size_t)b * nvc + j) * TC_D + t] = make_float2(y[t], y[t + TC_D]);
    if (t == 0 && rnv) rnv[(size_t)b * nvc + j] = rn;     // 1 / |C_{j-1} y_{j-1}|  (adjoint pass)
    // y_new[n] = sum_m P^T[m][n] y[m], the m range split over the two thread halves
    const float* PT = pt + (j & 1) * TC_N * TC_N;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
    for (int m = 64 * half; m < 64 * half + 64; m += 2) {
      acc0 = fmaf(PT[m * TC_N + n], y[m], acc0);
      acc1 = fmaf(PT[(m + 1) * TC_N + n], y[m + 1], acc1);
    }
    part[half * TC_N + n] = acc0 + acc1;
    __syncthreads();
    if (t < TC_N) y[t] = part[t] + part[TC_N + t];
    __syncthreads();
  }
  cp_async_wait<0>();
}

// Adjoint of the boundary pass.  With F_j = the loss of chunks j.. as a (scale-invariant) function of
// the start state y_j of chunk j, y_{j+1} = C_j y_j / |C_j y_j|:
//     Lam_j = d_j + C_j^dag Lam_{j+1} / |C_j y_j|,        Lam_nvc = 0,
// d_j = the adjoint of y_j from chunk j's own loss terms (virtual-clip backward run with a zero end
// adjoint).  Emits lamend[j] = Lam_{j+1}, the end adjoint the second virtual-clip backward starts
// chunk j from.  grid = B, block = 256 (8 warps x 16 rows of P^T, lanes across the columns).
__global__ void __launch_bounds__(256)
    psi_scan_boundary_bwd_kernel(const float* __restrict__ opsT, const float* __restrict__ rnv,
                                 const float2* __restrict__ dvec, int nvc, float2* __restrict__ lamend) {
  extern __shared__ __align__(16) unsigned char bsm_raw[];
  float* pt = reinterpret_cast<float*>(bsm_raw);          // [2][128][128]
  float* v = pt + 2 * TC_N * TC_N;                        // [128]  Lam_{j+1}, real form (Re | Im)
  const int t = threadIdx.x, b = blockIdx.x, lane = t & 31, warp = t >> 5;
  auto issue = [&](int j) {
    const float* src = opsT + ((size_t)b * nvc + j) * TC_N * TC_N;
    float* dst = pt + (j & 1) * TC_N * TC_N;
    for (int idx = t; idx < TC_N * TC_N / 4; idx += 256) cp_async16(dst + 4 * idx, src + 4 * idx);
  };
  if (t < TC_N) v[t] = 0.f;
  if (t < TC_D) lamend[((size_t)b * nvc + nvc - 1) * TC_D + t] = make_float2(0.f, 0.f);
  if (nvc >= 2) issue(nvc - 1);
  cp_async_commit();
  for (int j = nvc - 1; j >= 1; --j) {
    if (j - 1 >= 1) issue(j - 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();   // operator j landed; v complete
    const float* PT = pt + (j & 1) * TC_N * TC_N;
    const float sc = (j + 1 < nvc) ? rnv[(size_t)b * nvc + j + 1] : 0.f;
    float vr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) vr[r] = v[lane + 32 * r];
    float mine = 0.f;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const float* row = PT + (warp * 16 + rr) * TC_N;     // (C^dag lam)[m] = sum_n P^T[m][n] lam[n]
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) acc = fmaf(row[lane + 32 * r], vr[r], acc);
      acc = warp_sum_f(acc);
      if (rr == lane) mine = acc;
    }
    __syncthreads();   // every read of v done
    if (lane < 16) {
      const int m = warp * 16 + lane;
      const float2 d = dvec[((size_t)b * nvc + j) * TC_D + (m & (TC_D - 1))];
      v[m] = fmaf(sc, mine, m < TC_D ? d.x : d.y);
    }
    __syncthreads();
    if (t < TC_D) lamend[((size_t)b * nvc + j - 1) * TC_D + t] = make_float2(v[t], v[t + TC_D]);
  }
  cp_async_wait<0>();
}

// per-virtual-clip weights for the shared gradient epilogue: wv[b * nvc + j] = w[b]
__global__ void psi_scan_expand_w_kernel(const float* __restrict__ w, int B, int nvc, float* __restrict__ wv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < B * nvc) wv[v] = w[v / nvc];
}

// per-clip loss = sum over the clip's virtual clips
__global__ void psi_scan_sum_kernel(const double* __restrict__ lossv, int B, int nvc, float* __restrict__ loss,
                                    double* __restrict__ lossd) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    double s = 0.0;
    for (int j = 0; j < nvc; ++j) s += lossv[(size_t)b * nvc + j];
    loss[b] = (float)s;
    if (lossd) lossd[b] = s;
  }
}

}  // namespace amps
