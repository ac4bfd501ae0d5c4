// Shared device helpers for the AudioMPS scan kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amps {

// Steps per chunk.  The sequential chain runs CH steps between the batched (lane-parallel)
// reductions that produce E_k, nu_k^2, log terms and the rescale of the un-normalised state.
constexpr int CH = 32;

// ---- complex helpers on float2 (re, im) ---------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// conj(a) * b
__device__ __forceinline__ float2 cmul_ca(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
// acc += a * x
__device__ __forceinline__ void cmac(float2& acc, float2 a, float2 x) {
  acc.x = fmaf(a.x, x.x, acc.x);
  acc.y = fmaf(a.x, x.y, acc.y);
  acc.x = fmaf(-a.y, x.y, acc.x);
  acc.y = fmaf(a.y, x.x, acc.y);
}
// acc += a * conj(x)
__device__ __forceinline__ void cmac_cx(float2& acc, float2 a, float2 x) {
  acc.x = fmaf(a.x, x.x, acc.x);
  acc.y = fmaf(a.y, x.x, acc.y);
  acc.x = fmaf(a.y, x.y, acc.x);
  acc.y = fmaf(-a.x, x.y, acc.y);
}
__device__ __forceinline__ float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }

// ---- cp.async (LDGSTS) --------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- predicated shared-memory stores (keeps the per-step stores branch-free) ----------------
__device__ __forceinline__ void sts_if(bool on, float2* p, float2 v) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
  asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %0, 0;\n\t@pp st.shared.v2.f32 [%1], {%2, %3};\n\t}\n" ::"r"((int)on),
               "r"(a), "f"(v.x), "f"(v.y)
               : "memory");
}
__device__ __forceinline__ void sts_if(bool on, float* p, float v) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
  asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %0, 0;\n\t@pp st.shared.f32 [%1], %2;\n\t}\n" ::"r"((int)on), "r"(a),
               "f"(v)
               : "memory");
}

// 128-bit shared load as a volatile asm statement: volatile asms keep their relative order, so a run of
// these is issued back to back exactly where it is written (ptxas otherwise sinks each LDS next to
// its first use and exposes one shared-memory latency per load on the chain's critical path), without
// a dependent FADD chain tying the loads together, which an earlier version used (ncu source page: ~30 cycles per
// step, profiles/r1_chain_source.md).  `saddr` is a shared-window address (see smem_addr_pinned).
__device__ __forceinline__ float4 lds128v(unsigned saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// Shared-window address of `p`, made opaque to the optimiser: without this ptxas re-derives the
// per-thread address from %tid inside the step loop (an S2R + three ALU ops on the critical path
// after every second barrier).
__device__ __forceinline__ unsigned smem_addr_pinned(const void* p) {
  unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
  asm volatile("" : "+r"(a));
  return a;
}

// Address-based shared-memory accessors (32-bit shared-window addresses, see smem_addr_pinned): hot
// loops that index shared memory through generic pointers make the compiler convert generic ->
// shared inside the loop, which on sm_100 re-reads %cluster_ctaid (S2UR/S2R, tens of cycles).
__device__ __forceinline__ float lds32a(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64a(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32a_if(bool on, unsigned a, float v) {
  asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %0, 0;\n\t@pp st.shared.f32 [%1], %2;\n\t}\n" ::"r"((int)on), "r"(a), "f"(v)
               : "memory");
}
__device__ __forceinline__ void sts64a_if(bool on, unsigned a, float2 v) {
  asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %0, 0;\n\t@pp st.shared.v2.f32 [%1], {%2, %3};\n\t}\n" ::"r"((int)on), "r"(a),
               "f"(v.x), "f"(v.y)
               : "memory");
}

// The trainable scalar A (model.py:19) reaches the kernels either by value (host float) or through a
// device pointer, so that a training loop never has to read the parameter back to the host.
struct AVal {
  float host;
  const float* dev;
};
__device__ __forceinline__ float a_get(AVal a) { return a.dev ? __ldg(a.dev) : a.host; }

// Time-window ("segment") launch of the scan kernels: the checkpointed backward replays and
// differentiates one window of K steps at a time.  The host passes x / q_k / t_k pointers already
// offset to the window's first step; the kernels then only need the row stride of the waveform, the
// per-clip start states and end adjoints, and whether their gradient outputs continue a running sum.
struct SegFwd {
  int xstride;            // samples between consecutive clips of x (= the full clip length)
  const float2* x0;       // per-clip start state, x0[b * x0_stride + r] (null: the shared psi_0)
  int x0_stride;          // in float2 units
  float2* ckpt;           // checkpoint out: state at the start of every ck_chunks-th chunk (null: none)
  int ck_chunks;          // checkpoint interval in chunks
  int ck_stride;          // float2 units between consecutive clips' checkpoint rows
};
struct SegBwd {
  int xstride;
  const float2* lam_end;  // per-clip adjoint of the window's end state [B][DP] (null: zero)
  int accumulate;         // G / g_f / dA outputs continue the sums already in the output buffers
  int tprev_valid;        // t_{k-1} exists for the window's first step (window does not start at step 0)
};

// ---- thread map ---------------------------------------------------------------------------
// A CTA of DP*NQ threads owns one clip.  Thread t = i*NQ + jq holds, for matrix row i, the
// CPT = DP/NQ columns  col(c) = 2*NQ*(c/2) + 2*jq + (c&1)  in registers, so that for a fixed
// pair index the NQ lanes of a row group read 16*NQ contiguous bytes of the state vector
// (one conflict-free LDS.128 each).  Row sums are finished with xor-shuffles over the NQ lanes.
template <int DP, int NQ>
struct Map {
  static constexpr int NT = DP * NQ;
  static constexpr int CPT = DP / NQ;
  static constexpr int NP = CPT / 2;
  static_assert(DP % (2 * NQ) == 0, "DP must be a multiple of 2*NQ");
  static_assert(NT % 32 == 0, "CTA must be whole warps");
  static_assert(NQ == 4 || NQ == 8 || NQ == 16, "NQ in {4,8,16}");
  __device__ static __forceinline__ int col(int c, int jq) {
    return 2 * NQ * (c >> 1) + 2 * jq + (c & 1);
  }
};

template <int NQ>
__device__ __forceinline__ float2 group_sum(float2 v) {
#pragma unroll
  for (int m = 1; m < NQ; m <<= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, m);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, m);
  }
  return v;
}

// Row-sum of a COMPLEX partial over the NQ lanes with half the shuffles: after the first level the even
// lanes carry only the real part and the odd lanes only the imaginary part (SHFL issues at one warp
// instruction per cycle per SM, so with 16 warps per SM every shuffle per thread costs 16 cycles per
// step).  Returns Re(sum) on even lanes, Im(sum) on odd lanes.
template <int NQ>
__device__ __forceinline__ float pair_reduce(float2 v, int jq) {
  const bool odd = jq & 1;
  float r = (odd ? v.y : v.x) + __shfl_xor_sync(0xffffffffu, odd ? v.x : v.y, 1);
#pragma unroll
  for (int m = 2; m < NQ; m <<= 1) r += __shfl_xor_sync(0xffffffffu, r, m);
  return r;
}

// Row-sum over the NQ lanes of a group on the chain's critical path.  For NQ = 4 the three partner
// values are fetched with three INDEPENDENT shuffles per component (one shuffle latency, ~29 cycles on
// B200) instead of two dependent butterfly levels (two latencies); the summation tree
// (own + xor1) + (xor2 + xor3) is the same on every lane, so the replicated result is bit-identical.
template <int NQ>
__device__ __forceinline__ float2 group_sum_fast(float2 v) {
  if (NQ == 4) {
    const float x1 = __shfl_xor_sync(0xffffffffu, v.x, 1), y1 = __shfl_xor_sync(0xffffffffu, v.y, 1);
    const float x2 = __shfl_xor_sync(0xffffffffu, v.x, 2), y2 = __shfl_xor_sync(0xffffffffu, v.y, 2);
    const float x3 = __shfl_xor_sync(0xffffffffu, v.x, 3), y3 = __shfl_xor_sync(0xffffffffu, v.y, 3);
    return make_float2((v.x + x1) + (x2 + x3), (v.y + y1) + (y2 + y3));
  }
  return group_sum<NQ>(v);
}

// Load this thread's register slice of row i of a [DP][DP] complex matrix.
template <int DP, int NQ>
__device__ __forceinline__ void load_slice(float2 (&dst)[DP / NQ], const float2* __restrict__ mat,
                                           int i, int jq) {
#pragma unroll
  for (int c = 0; c < DP / NQ; ++c) dst[c] = mat[i * DP + Map<DP, NQ>::col(c, jq)];
}

// Partial complex mat-vec of two register slices against one smem vector:
//   a += sum_c A[c] * v[col(c)],  b += sum_c Bm[c] * v[col(c)]
template <int DP, int NQ>
__device__ __forceinline__ void matvec2(const float2 (&A)[DP / NQ], const float2 (&Bm)[DP / NQ],
                                        const float2* __restrict__ v, int jq, float2& a,
                                        float2& b) {
  float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
  for (int m = 0; m < DP / NQ / 2; ++m) {
    const float4 xv = *reinterpret_cast<const float4*>(&v[2 * NQ * m + 2 * jq]);
    const float2 x0 = make_float2(xv.x, xv.y), x1 = make_float2(xv.z, xv.w);
    cmac(a0, A[2 * m], x0);
    cmac(a1, A[2 * m + 1], x1);
    cmac(b0, Bm[2 * m], x0);
    cmac(b1, Bm[2 * m + 1], x1);
  }
  a = make_float2(a0.x + a1.x, a0.y + a1.y);
  b = make_float2(b0.x + b1.x, b0.y + b1.y);
}

template <int DP, int NQ>
__device__ __forceinline__ float2 matvec1(const float2 (&A)[DP / NQ], const float2* __restrict__ v,
                                          int jq) {
  float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
  for (int m = 0; m < DP / NQ / 2; ++m) {
    const float4 xv = *reinterpret_cast<const float4*>(&v[2 * NQ * m + 2 * jq]);
    cmac(a0, A[2 * m], make_float2(xv.x, xv.y));
    cmac(a1, A[2 * m + 1], make_float2(xv.z, xv.w));
  }
  return make_float2(a0.x + a1.x, a0.y + a1.y);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

}  // namespace amps

// ---- thread-block cluster / distributed shared memory helpers --------------------------------
namespace amps {
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory object in CTA `rank` of this cluster
__device__ __forceinline__ unsigned dsmem_addr(const void* local, unsigned rank) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(local));
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(unsigned addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_dsmem_f4(unsigned addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_dsmem_f1(unsigned addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
// predicated 8-byte store into another CTA's shared memory (no divergent branch on the chain)
__device__ __forceinline__ void st_dsmem_f2_if(bool on, unsigned addr, float2 v) {
  asm volatile(
      "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %0, 0;\n\t@pp st.shared::cluster.v2.f32 [%1], {%2, %3};\n\t}\n" ::"r"((int)on),
      "r"(addr), "f"(v.x), "f"(v.y)
      : "memory");
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// full cluster barrier: every thread of every CTA; orders shared::cta and shared::cluster accesses
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive_release();
  cluster_wait_acquire();
}
}  // namespace amps

// ---- mbarrier hand-offs between the two CTAs of a cluster --------------------------------------
namespace amps {
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init_cluster() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
// local arrive that also announces `bytes` of st.async traffic for the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
// predicated 8-byte store into (possibly another) CTA's shared memory that completes `8` bytes on the
// mbarrier `mbar_addr` of the SAME target CTA (both shared::cluster addresses): data and signal in one
// instruction, no release fence / cluster barrier on the sender's critical path.
__device__ __forceinline__ void st_async_f2_if(bool on, unsigned addr, float2 v, unsigned mbar_addr) {
  asm volatile(
      "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %0, 0;\n\t"
      "@pp st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%1], {%2, %3}, [%4];\n\t}\n" ::"r"((int)on),
      "r"(addr), "f"(v.x), "f"(v.y), "r"(mbar_addr)
      : "memory");
}
// arrive (release, cluster scope) on an mbarrier that lives in another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_remote(unsigned remote_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(remote_addr) : "memory");
}
// block until the phase with the given parity of a LOCAL mbarrier has completed (acquire, cluster)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
// Same wait with the default (CTA-scope) acquire: for phases completed by st.async ... complete_tx the
// data is made visible to the waiters by the mbarrier itself, and the cluster-scope acquire above
// costs a CCTL.IVALL (L1 invalidate) after EVERY wait -- 20 % of the D = 128 forward's step loop on the
// ncu source page.
__device__ __forceinline__ void mbar_wait_cta(unsigned long long* bar, unsigned parity) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
}  // namespace amps
