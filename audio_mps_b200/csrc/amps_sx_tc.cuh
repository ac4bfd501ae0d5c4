// Expectation pass of the forward on the tensor cores (D = 33..64).
//
// Per step the forward needs, besides the chain x_{k+1} = c_k q_k (N + s_k R) x_k, the vector S x'_k
// (S = R + R^dag; kept for the adjoint sweep), E_k = x'_k^dag S x'_k / |x_k|^2 and the loss term
// -log1p(E_k inc_k / A) (model.py:293-294, 319-325).  None of that feeds the next state.  Inside the
// sequential kernel it was a second mat-vec per step (32 FFMA + 4 LDS.128 per thread in the chain's
// shuffle shadows); over the time axis it is ONE GEMM,  [S x'_k]_k = S_real X'^T  (2D x 2D times 2D x T),
// fully parallel.  The sequential kernel now stores x'_k (where S x'_k used to go) and |x_k|^2, and this
// kernel transforms the trajectory IN PLACE:
//     rows of sptraj:  x'_k  ->  S x'_k            ev[k] = (E_k, |x_k|^2)            loss_part[b][split]
// A = S in real form (row 2i+c, column 2j+c'; [[Sr, -Si], [Si, Sr]]), tf32 hi + lo, staged ONCE per CTA
// (128 KB); B = 32 steps of x' per stage (row = step: the trajectory's own layout IS K-major), tf32 hi + lo;
// three passes S_hi X_hi + S_lo X_hi + S_hi X_lo (2^-21 per product: E_k is a small residual of large terms,
// a single rounded x' operand was worth 1e-4 of a 32-step loss).  Accumulators: two 32-column buffers in
// tensor memory; warp 16 issues the MMAs; the 16 worker warps stage tile j+1, then drain tile j:
// tcgen05.ld -> shared transpose -> coalesced read-modify-write of the trajectory rows with the E_k dot
// products folded in.
#pragma once
#include "amps_common.cuh"
#include "amps_scan_tc.cuh"
#include "amps_tiles_tc.cuh"

namespace amps {

constexpr int SX_NS = 32;             // steps per tile = UMMA N (D = 128 kernel)
constexpr int SX_THREADS = 512;       // worker threads (warps 0..15)
constexpr int SX_BLOCK = SX_THREADS + 32;
constexpr int SX1_NS = 64;            // steps per tile of the D = 64 kernel

struct SxArgs {
  const float2* matS;      // [DP][DP]
  float2* sptraj;          // [B][T][DP]  in: x'_k (row k), out: S x'_k
  float2* ev;              // [B][T]      in: (., |x_k|^2), out: (E_k, |x_k|^2)
  const float* x;          // waveform, clip stride xstride
  double* loss_part;       // [B][nsplit]
  const float4* spanel;    // D = 128: [8 phases][8][512] panel-ordered tf32 hi/lo of S (psi_sx2_panel_kernel)
  int T, xstride, nsplit, steps_per_split;   // steps_per_split: multiple of 32
  AVal A;
};

// D = 64.  A = S (real form, tf32 hi and lo) lives in TENSOR MEMORY for the whole kernel (TS-mode MMA: columns
// [0,128) = S_hi, [128,256) = S_lo, written once with tcgen05.st), so an MMA reads only the 64-row x' tile from
// shared memory: with both operands in shared memory every K = 8 instruction re-read a 128 x 8 slice of S (32
// wavefronts) for 17 cycles of math, and the kernel sat at 15-20 % tensor activity.  Two 64-column
// accumulators ([256,320), [320,384)); two 64 KB stages of x' (hi, lo).
template <int DP>
struct alignas(1024) SxSmem {
  static constexpr int KR = 2 * DP;                 // real contraction length
  static constexpr int NKB = KR / 32;               // 128-byte K blocks
  uint8_t b[2][NKB][SX1_NS * TL_ROWB];              // x' tile truncated to tf32, row = step
  uint8_t b_lo[2][NKB][SX1_NS * TL_ROWB];           // x' - trunc(x')
  float outs[SX1_NS][KR + 4];                       // (S x')[step][2i+c], epilogue transpose
  double lred[16];
  unsigned long long full_bar[2], empty_bar[2], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// grid = B * nsplit, block = 544
template <int DP>
__global__ void __launch_bounds__(SX_BLOCK, 1) psi_sx_tc_kernel(SxArgs g) {
  using Sm = SxSmem<DP>;
  constexpr int KR = Sm::KR, NKB = Sm::NKB, NS = SX1_NS;
  static_assert(DP == 64, "one UMMA M = 128 tile of output rows");
  constexpr uint32_t COL_AH = 0, COL_AL = 128, COL_ACC = 256;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem_al = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  Sm& sm = *reinterpret_cast<Sm*>(smem_al);
  const float A = a_get(g.A);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / g.nsplit, sp = blockIdx.x % g.nsplit;
  const int nsteps = g.T - 1;
  const int k_begin = sp * g.steps_per_split;
  const int nloc = max(0, min(g.steps_per_split, nsteps - k_begin));
  const int ntile = (nloc + NS - 1) / NS;
  float* rows = reinterpret_cast<float*>(g.sptraj + ((size_t)b * g.T + k_begin) * DP);   // [step][KR] floats
  float2* evb = g.ev + (size_t)b * g.T + k_begin;
  const float* xb = g.x + (size_t)b * g.xstride + k_begin;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.full_bar[s], SX_THREADS);
      mbar_init(&sm.empty_bar[s], 1);
      mbar_init(&sm.acc_full[s], 1);
      mbar_init(&sm.acc_empty[s], SX_THREADS);
    }
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
  const int q4 = warp & 3, cg = warp >> 2;          // (worker warps) TMEM lane quarter, column group
  // A = S real form: lane m = 2i + c holds row m; warp (q4, cg) writes contraction columns [32 cg, 32 cg + 32),
  // i.e. complex columns j = 16 cg .. 16 cg + 15 of row i:  Re-out row: (Sr, -Si) pairs, Im-out row: (Si, Sr)
  if (warp < SX_THREADS / 32) {
    const int m = 32 * q4 + lane, i = m >> 1, c = m & 1;
    float hi[32], lo[32];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const float2 sv = g.matS[i * DP + 16 * cg + jj];
      const float v0 = c ? sv.y : sv.x, v1 = c ? sv.x : -sv.y;
      hi[2 * jj] = tc_trunc_tf32(v0);
      hi[2 * jj + 1] = tc_trunc_tf32(v1);
      lo[2 * jj] = v0 - hi[2 * jj];
      lo[2 * jj + 1] = v1 - hi[2 * jj + 1];
    }
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q4) << 16);
    tc_st32(lane_base + COL_AH + 32 * cg, hi);
    tc_st32(lane_base + COL_AL + 32 * cg, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 64, M = 128
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  if (warp == SX_THREADS / 32) {
    // ---- MMA warp -----------------------------------------------------------------------------
    if (lane == 0) {
      for (int j = 0; j < ntile; ++j) {
        const int s = j & 1;
        mbar_wait_cta(&sm.full_bar[s], (j >> 1) & 1);
        if (j >= 2) mbar_wait_cta(&sm.acc_empty[s], ((j >> 1) - 1) & 1);   // tile j-2 drained out of accumulator s
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        bool first = true;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {      // S_hi X_hi + S_lo X_hi + S_hi X_lo
          const uint32_t acol = tmem + (pass == 1 ? COL_AL : COL_AH);
#pragma unroll 1
          for (int kb = 0; kb < NKB; ++kb) {
            const uint64_t db0 = tc_make_desc(tc_smem_u32(pass == 2 ? sm.b_lo[s][kb] : sm.b[s][kb]));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              tc_mma_ts(tmem + COL_ACC + NS * s, acol + kb * 32 + ks * 8, db0 + 2 * ks, idesc, first ? 0u : 1u);
              first = false;
            }
          }
        }
        tl_commit(&sm.empty_bar[s]);
        tl_commit(&sm.acc_full[s]);
      }
    }
  } else {
    // ---- worker warps ---------------------------------------------------------------------------
    // A tile is 64 rows (steps) x KR floats; warp w owns rows w + 16 q, lane l the 16-byte chunk l of the row
    // (one coalesced 512-byte row per warp instruction).  The same map serves the staging AND the drain, so
    // the x' values a thread staged are the ones it needs for the E_k dot product two iterations later:
    // three register sets roll through fetch(j+1) / stage(j) / drain(j-1) and nothing is read twice.
    double lossacc = 0.0;
    constexpr int NCH = NS / 16;           // rows per warp
    static_assert(KR / 4 == 32, "one lane per 16-byte chunk of a 128-float row");
    struct Pre {
      float4 v[NCH];
    };
    auto fetch = [&](Pre& P, int j) {
      const int n0 = j * NS, len = min(NS, nloc - n0);
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        const int n = warp + 16 * q;
        P.v[q] = (n < len) ? *reinterpret_cast<const float4*>(rows + (size_t)(n0 + n) * KR + 4 * lane)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto stage = [&](const Pre& P, int j) {
      const int s = j & 1;
      if (j >= 2) mbar_wait_cta(&sm.empty_bar[s], ((j >> 1) - 1) & 1);
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        const int n = warp + 16 * q;
        const int kb = lane / 8, ch = lane % 8;
        const float4 v = P.v[q];
        const float4 h = make_float4(tc_trunc_tf32(v.x), tc_trunc_tf32(v.y), tc_trunc_tf32(v.z), tc_trunc_tf32(v.w));
        *reinterpret_cast<float4*>(sm.b[s][kb] + tl_off(n, ch)) = h;
        *reinterpret_cast<float4*>(sm.b_lo[s][kb] + tl_off(n, ch)) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      tl_mbar_arrive(&sm.full_bar[s]);
    };
    // drain tile j: accumulator -> outs (transposed) -> rows (in place), E_k, loss
    auto drain = [&](const Pre& P, int j) {
      const int s = j & 1;
      const int n0 = j * NS, len = min(NS, nloc - n0);
      float nu2[NCH], inc[NCH];
#pragma unroll
      for (int q = 0; q < NCH; ++q) {       // the step's scalars, requested before the accumulator is awaited
        const int n = warp + 16 * q;
        nu2[q] = 1.f;
        inc[q] = 0.f;
        if (lane == 0 && n < len) {
          nu2[q] = evb[n0 + n].y;
          inc[q] = xb[n0 + n + 1] - xb[n0 + n];
        }
      }
      mbar_wait_cta(&sm.acc_full[s], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      {
        // warp (q4, cg): lanes 32 q4 .. (output rows m), steps 16 cg .. 16 cg + 15 of the tile
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(32 * q4) << 16) + COL_ACC + NS * s + 16 * cg;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const int m = 32 * q4 + lane;
#pragma unroll
        for (int e = 0; e < 16; ++e) sm.outs[16 * cg + e][m] = __uint_as_float(r[e]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      tl_mbar_arrive(&sm.acc_empty[s]);
      bar_named(5, SX_THREADS);
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        const int n = warp + 16 * q;
        float e = 0.f;
        if (n < len) {
          const float4 o = *reinterpret_cast<const float4*>(&sm.outs[n][4 * lane]);
          const float4 x = P.v[q];
          e = x.x * o.x + x.y * o.y + x.z * o.z + x.w * o.w;
          *reinterpret_cast<float4*>(rows + (size_t)(n0 + n) * KR + 4 * lane) = o;
        }
        e = warp_sum_f(e);
        if (lane == 0 && n < len) {
          const float E = e / fmaxf(nu2[q], 1e-12f);                    // model.py:324-325 on x'
          lossacc -= (double)log1pf((E * inc[q]) / A);                  // model.py:294
          evb[n0 + n] = make_float2(E, nu2[q]);
        }
      }
      bar_named(5, SX_THREADS);   // outs free for the next tile
    };
    {
      Pre P0, P1, P2;
      auto iter = [&](Pre& cur, Pre& nxt, Pre& prv, int j) {   // j < ntile
        if (j + 1 < ntile) fetch(nxt, j + 1);
        stage(cur, j);
        if (j > 0) drain(prv, j - 1);
      };
      if (ntile > 0) fetch(P0, 0);
      for (int j = 0; j < ntile; j += 3) {
        iter(P0, P1, P2, j);
        if (j + 1 < ntile) iter(P1, P2, P0, j + 1);
        if (j + 2 < ntile) iter(P2, P0, P1, j + 2);
      }
      if (ntile > 0) {
        const int last = ntile - 1;
        if (last % 3 == 0) drain(P0, last);
        else if (last % 3 == 1) drain(P1, last);
        else drain(P2, last);
      }
    }
    lossacc = warp_sum_d(lossacc);
    if (lane == 0) sm.lred[warp] = lossacc;
    bar_named(5, SX_THREADS);
    if (tid == 0) {
      double tot = 0.0;
      for (int wv = 0; wv < SX_THREADS / 32; ++wv) tot += sm.lred[wv];
      g.loss_part[blockIdx.x] = tot;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

// -------------------------------------------------------------------------------------------
// D = 65..128: S in real form is 256 x 256 (512 KB with its lo part) -- it fits neither tensor nor shared memory.
// The CTA walks its steps in BLOCKS of 128 (2 tiles of 64 steps) and, per block, through eight phases
// (contraction quarter kq, output-row half mh): a phase writes the 128 x 64 panel S[mh][kq] (hi + lo, 128 columns)
// into one of TWO tensor-memory buffers (columns [0,256); TS-mode MMA as in the D = 64 kernel), so the panel of
// phase p+1 is written while the MMAs of phase p run.  The x' tiles of a quarter (64 steps x 64 floats, hi + lo) are
// staged once and used by both halves.  The 4 accumulators (2 halves x 2 tiles x 64 columns) live in columns
// [256,512); only after the eighth phase are the rows rewritten (x' -> S x'), so the in-place update never
// overtakes a read.
// -------------------------------------------------------------------------------------------
constexpr int SX2_TPB = 2;             // tiles per block
constexpr int SX2_NS = 64;             // steps per tile
constexpr int SX2_SLOTS = 4;
struct alignas(1024) Sx2Smem {
  uint8_t b[SX2_SLOTS][2][SX2_NS * TL_ROWB];   // x' tile, one contraction quarter (2 K blocks of 32), hi
  uint8_t b_lo[SX2_SLOTS][2][SX2_NS * TL_ROWB];
  float outs[SX2_NS][128 + 4];
  float esum[SX2_TPB * SX2_NS];        // sum_i Re(conj(x'_i) (S x')_i) per step of the block
  double lred[16];
  unsigned long long a_full[2], a_empty[2], b_full[SX2_SLOTS], b_empty[SX2_SLOTS], acc_full, acc_empty;
  uint32_t tmem_base;
};

__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}

// Panel table of S for psi_sx2_tc_kernel: phase p = 2 kq + mh, worker thread t = 32 warp + lane with (q4, cg) =
// (warp & 3, warp >> 2).  TMEM lane m = 32 q4 + lane = 2 il + c holds real row c of complex output row
// i = 64 mh + il; the thread owns panel columns [16 cg, 16 cg + 16) = complex columns j = 32 kq + 8 cg + (0..7)
// (real form: row Re = (Re S, -Im S), row Im = (Im S, Re S)).  Stored as float4 e (0..3 = hi, 4..7 = lo) at
// [(8 p + e) * 512 + t], so a warp's load is one contiguous 512 bytes (reading S itself here would be 32
// different rows per instruction -- that serialised the L1 tag stage and was 60 % of the kernel's time).
// grid = 8, block = 512
__global__ void __launch_bounds__(SX_THREADS) psi_sx2_panel_kernel(const float2* __restrict__ matS, float4* __restrict__ spanel) {
  constexpr int DP = 128;
  const int p = blockIdx.x, kq = p >> 1, mh = p & 1;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q4 = warp & 3, cg = warp >> 2;
  const int m = 32 * q4 + lane, c = m & 1;
  const float2* src = matS + (size_t)(64 * mh + (m >> 1)) * DP + 32 * kq + 8 * cg;
  float hi[16], lo[16];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const float2 sv = src[jj];
    const float v0 = c ? sv.y : sv.x, v1 = c ? sv.x : -sv.y;
    hi[2 * jj] = tc_trunc_tf32(v0);
    hi[2 * jj + 1] = tc_trunc_tf32(v1);
    lo[2 * jj] = v0 - hi[2 * jj];
    lo[2 * jj + 1] = v1 - hi[2 * jj + 1];
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    spanel[(size_t)(8 * p + e) * SX_THREADS + t] = make_float4(hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
    spanel[(size_t)(8 * p + 4 + e) * SX_THREADS + t] = make_float4(lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
  }
}

// grid = B * nsplit, block = 544
__global__ void __launch_bounds__(SX_BLOCK, 1) psi_sx2_tc_kernel(SxArgs g) {
  constexpr int DP = 128, KR = 256, NS = SX2_NS, BLK = SX2_TPB * NS;
  constexpr uint32_t COL_ACC = 256;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem_al = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
  Sx2Smem& sm = *reinterpret_cast<Sx2Smem*>(smem_al);
  const float A = a_get(g.A);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / g.nsplit, sp = blockIdx.x % g.nsplit;
  const int nsteps = g.T - 1;
  const int k_begin = sp * g.steps_per_split;
  const int nloc = max(0, min(g.steps_per_split, nsteps - k_begin));
  const int nblk = (nloc + BLK - 1) / BLK;
  float* rows = reinterpret_cast<float*>(g.sptraj + ((size_t)b * g.T + k_begin) * DP);
  float2* evb = g.ev + (size_t)b * g.T + k_begin;
  const float* xb = g.x + (size_t)b * g.xstride + k_begin;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.a_full[s], SX_THREADS);
      mbar_init(&sm.a_empty[s], 1);
    }
    for (int s = 0; s < SX2_SLOTS; ++s) {
      mbar_init(&sm.b_full[s], SX_THREADS);
      mbar_init(&sm.b_empty[s], 1);
    }
    mbar_init(&sm.acc_full, 1);
    mbar_init(&sm.acc_empty, SX_THREADS);
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
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  auto ntile_of = [&](int blk) { return (min(BLK, nloc - blk * BLK) + NS - 1) / NS; };

  if (warp == SX_THREADS / 32) {
    // ---- MMA warp -----------------------------------------------------------------------------
    if (lane == 0) {
      int gp = 0, jb = 0;      // phases and staged tiles so far (barrier parities)
      for (int blk = 0; blk < nblk; ++blk) {
        const int nt = ntile_of(blk);
        if (blk > 0) mbar_wait_cta(&sm.acc_empty, (blk - 1) & 1);     // the previous block is drained
        for (int kq = 0; kq < 4; ++kq, jb += nt)
          for (int mh = 0; mh < 2; ++mh, ++gp) {
            const int ab = gp & 1;
            mbar_wait_cta(&sm.a_full[ab], (gp >> 1) & 1);
            for (int tl = 0; tl < nt; ++tl) {
              const int j = jb + tl, s = j % SX2_SLOTS;
              if (mh == 0) mbar_wait_cta(&sm.b_full[s], (j / SX2_SLOTS) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
              bool first = kq == 0;
#pragma unroll 1
              for (int pass = 0; pass < 3; ++pass) {      // S_hi X_hi + S_lo X_hi + S_hi X_lo
                const uint32_t acol = tmem + 128 * ab + (pass == 1 ? 64 : 0);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                  const uint64_t db0 = tc_make_desc(tc_smem_u32(pass == 2 ? sm.b_lo[s][kb] : sm.b[s][kb]));
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    tc_mma_ts(tmem + COL_ACC + (uint32_t)((mh * SX2_TPB + tl) * NS), acol + kb * 32 + ks * 8, db0 + 2 * ks,
                              idesc, first ? 0u : 1u);
                    first = false;
                  }
                }
              }
              if (mh == 1) tl_commit(&sm.b_empty[s]);
            }
            tl_commit(&sm.a_empty[ab]);
          }
        tl_commit(&sm.acc_full);
      }
    }
  } else {
    // ---- worker warps ---------------------------------------------------------------------------
    double lossacc = 0.0;
    const int q4 = warp & 3, cg = warp >> 2;          // TMEM lane quarter, column group
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q4) << 16);
    const int srow = tid >> 4, sc = tid & 15;         // staging map: rows srow, srow + 32; 16-byte chunk sc of the quarter
    // panel S[mh][kq] comes pre-split and in thread order from psi_sx2_panel_kernel: 8 coalesced float4 per phase
    float4 apre[8];
    auto fetch_a = [&](int kq, int mh) {
      const float4* src = g.spanel + (size_t)(8 * (2 * kq + mh)) * SX_THREADS + tid;
#pragma unroll
      for (int e = 0; e < 8; ++e) apre[e] = __ldg(src + (size_t)e * SX_THREADS);
    };
    float4 pre[SX2_TPB][2];
    auto fetch_b = [&](int n0b, int kq) {
#pragma unroll
      for (int tl = 0; tl < SX2_TPB; ++tl) {
        const int n0 = n0b + tl * NS, len = min(NS, nloc - n0);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int n = srow + 32 * q;
          pre[tl][q] = (n < len) ? *reinterpret_cast<const float4*>(rows + (size_t)(n0 + n) * KR + 64 * kq + 4 * sc)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    int gp = 0, jb = 0;
    if (nblk > 0) {
      fetch_b(0, 0);
      fetch_a(0, 0);
    }
    for (int blk = 0; blk < nblk; ++blk) {
      const int n0b = blk * BLK, nt = ntile_of(blk);
      for (int kq = 0; kq < 4; ++kq) {
        // x' tiles of this quarter -> shared memory (rows fetched one quarter ahead)
#pragma unroll
        for (int tl = 0; tl < SX2_TPB; ++tl) {
          if (tl < nt) {
            const int j = jb + tl, s = j % SX2_SLOTS;
            if (j >= SX2_SLOTS) mbar_wait_cta(&sm.b_empty[s], ((j / SX2_SLOTS) - 1) & 1);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int n = srow + 32 * q, kb = sc >> 3, ch = sc & 7;
              const float4 v = pre[tl][q];
              const float4 h = make_float4(tc_trunc_tf32(v.x), tc_trunc_tf32(v.y), tc_trunc_tf32(v.z), tc_trunc_tf32(v.w));
              *reinterpret_cast<float4*>(sm.b[s][kb] + tl_off(n, ch)) = h;
              *reinterpret_cast<float4*>(sm.b_lo[s][kb] + tl_off(n, ch)) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            tl_mbar_arrive(&sm.b_full[s]);
          }
        }
        jb += nt;
        if (kq < 3) fetch_b(n0b, kq + 1);
        else if (blk + 1 < nblk) fetch_b(n0b + BLK, 0);
        // the two panels of this quarter
#pragma unroll
        for (int mh = 0; mh < 2; ++mh, ++gp) {
          float hi[16], lo[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[4 * e] = apre[e].x, hi[4 * e + 1] = apre[e].y, hi[4 * e + 2] = apre[e].z, hi[4 * e + 3] = apre[e].w;
            lo[4 * e] = apre[4 + e].x, lo[4 * e + 1] = apre[4 + e].y, lo[4 * e + 2] = apre[4 + e].z, lo[4 * e + 3] = apre[4 + e].w;
          }
          if (mh == 0) fetch_a(kq, 1);
          else fetch_a((kq + 1) & 3, 0);
          const int ab = gp & 1;
          if (gp >= 2) mbar_wait_cta(&sm.a_empty[ab], ((gp >> 1) - 1) & 1);   // the MMAs of phase gp - 2 have read it
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          tc_st16(lane_base + 128 * ab + 16 * cg, hi);
          tc_st16(lane_base + 128 * ab + 64 + 16 * cg, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          tl_mbar_arrive(&sm.a_full[ab]);
        }
      }
      // ---- drain the block: 2 halves x nt tiles; warp w owns rows w + 16 q, lane l the 16-byte chunk l ----
      mbar_wait_cta(&sm.acc_full, blk & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      for (int mh = 0; mh < 2; ++mh)
        for (int tl = 0; tl < nt; ++tl) {
          const int n0 = n0b + tl * NS, len = min(NS, nloc - n0);
          float4 xr[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {      // x' of this half, before it is overwritten (L2)
            const int n = warp + 16 * q;
            xr[q] = (n < len) ? *reinterpret_cast<const float4*>(rows + (size_t)(n0 + n) * KR + 128 * mh + 4 * lane)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          {
            uint32_t r[16];
            const uint32_t taddr = lane_base + COL_ACC + (uint32_t)((mh * SX2_TPB + tl) * NS + 16 * cg);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            const int m = 32 * q4 + lane;
#pragma unroll
            for (int e = 0; e < 16; ++e) sm.outs[16 * cg + e][m] = __uint_as_float(r[e]);
          }
          bar_named(5, SX_THREADS);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = warp + 16 * q;
            float e = 0.f;
            if (n < len) {
              const float4 o = *reinterpret_cast<const float4*>(&sm.outs[n][4 * lane]);
              e = xr[q].x * o.x + xr[q].y * o.y + xr[q].z * o.z + xr[q].w * o.w;
              *reinterpret_cast<float4*>(rows + (size_t)(n0 + n) * KR + 128 * mh + 4 * lane) = o;
            }
            e = warp_sum_f(e);
            if (lane == 0 && n < len) sm.esum[tl * NS + n] = (mh == 0 ? 0.f : sm.esum[tl * NS + n]) + e;
          }
          bar_named(5, SX_THREADS);
        }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      tl_mbar_arrive(&sm.acc_empty);      // every accumulator of the block has been read
      {
        const int nb = min(BLK, nloc - n0b);
        if (tid < nb) {
          const float nu2 = evb[n0b + tid].y;
          const float E = sm.esum[tid] / fmaxf(nu2, 1e-12f);                // model.py:324-325 on x'
          const float inc = xb[n0b + tid + 1] - xb[n0b + tid];
          lossacc -= (double)log1pf((E * inc) / A);                         // model.py:294
          evb[n0b + tid] = make_float2(E, nu2);
        }
      }
      bar_named(5, SX_THREADS);
    }
    lossacc = warp_sum_d(lossacc);
    if (lane == 0) sm.lred[warp] = lossacc;
    bar_named(5, SX_THREADS);
    if (tid == 0) {
      double tot = 0.0;
      for (int wv = 0; wv < SX_THREADS / 32; ++wv) tot += sm.lred[wv];
      g.loss_part[blockIdx.x] = tot;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

}  // namespace amps
