// Latency of ONE sequential "batched step" on the tensor cores: how fast can a CTA advance a
// recursion  X_{k+1} = f(A * X_k)  whose every step is a small UMMA (M = 128, K = 128, N = clips) that
// depends on the previous one?  Per step: thread 0 issues passes * K/8 tcgen05.mma (SS mode) and a
// commit; the 128 epilogue threads wait on the mbarrier, read their accumulator row from tensor
// memory, and write the next B operand (transposed, K-major) to shared memory; fence + barrier.
// This is the floor of the large-batch tensor-core step kernel discussed in DESIGN.md 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_step_latency tc_step_latency.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int M = 128, K = 128, KB = 32, NKB = K / KB;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

template <int N>
__global__ void __launch_bounds__(128) step_kernel(int iters, int passes, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                          // NKB tiles of 128 rows x 128 B
  uint8_t* sB = smem + NKB * M * 128;          // NKB tiles of N rows x 128 B
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < NKB * M * 32; i += 128) reinterpret_cast<float*>(sA)[i] = (i % 97 == 0) ? 0.01f : 0.f;
  for (int i = tid; i < NKB * N * 32; i += 128) reinterpret_cast<float*>(sB)[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  float acc_sink = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (tid == 0) {
      for (int p = 0; p < passes; ++p)
        for (int kb = 0; kb < NKB; ++kb) {
          const uint64_t da = make_desc(smem_u32(sA + kb * M * 128));
          const uint64_t db = make_desc(smem_u32(sB + kb * N * 128));
#pragma unroll
          for (int ks = 0; ks < KB / 8; ++ks) mma_tf32_ss(tmem, da + 2 * ks, db + 2 * ks, idesc, (p | kb | ks) ? 1u : 0u);
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar))
                   : "memory");
    }
    mbar_wait(&mbar, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // accumulator row `tid` (N columns) -> next B operand, transposed: B[n][k = tid], K-major swizzled
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      tmem_ld8(lane_base + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = c0 + j, kb = tid / KB, kc = tid % KB;
        float* dst = reinterpret_cast<float*>(sB + kb * N * 128 + n * 128 + (((kc / 4) ^ (n & 7)) * 16)) + (kc & 3);
        const float nv = v[j] * 0.5f + 0.25f;
        *dst = nv;
        acc_sink += nv;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  const long long t1 = clock64();
  if (tid == 0) *cycles = t1 - t0;
  sink[tid] = acc_sink;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem));
}

template <int N>
void run(int passes) {
  long long* d_c;
  float* d_s;
  cudaMalloc(&d_c, 8);
  cudaMalloc(&d_s, 128 * 4);
  const size_t smem = NKB * M * 128 + NKB * N * 128 + 1024;
  cudaFuncSetAttribute(step_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  step_kernel<N><<<1, 128, smem>>>(iters, passes, d_c, d_s);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d passes=%d: %s  %.0f cycles/step  (%d MMAs per step)\n", N, passes, cudaGetErrorString(e),
         (double)c / iters, passes * K / 8);
  cudaFree(d_c);
  cudaFree(d_s);
}

int main() {
  for (int passes = 1; passes <= 3; passes += 2) {
    run<8>(passes);
    run<16>(passes);
    run<32>(passes);
    run<64>(passes);
    run<128>(passes);
  }
  return 0;
}
