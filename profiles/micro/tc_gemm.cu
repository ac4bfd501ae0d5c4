// Stage-1 bring-up of the tcgen05 path: one CTA computes D[128x128] = A[128x128] * B[128x128]^T with
// tcgen05.mma kind::tf32 (operands K-major in shared memory, 128-byte swizzle, accumulator in TMEM),
// once in plain TF32 and once as the 3-pass split (A_hi B_hi + A_lo B_hi + A_hi B_lo) that recovers
// ~fp32 accuracy.  Verified against a double-precision host product.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 128, K = 64;   // 4 operand tiles (A, A_lo, B, B_lo) x 32 KB fit in shared memory
constexpr int KB = 32;                 // floats per 128-byte swizzle row
constexpr int NKB = K / KB;            // K blocks
constexpr int TILE_BYTES = M * 128;    // one K block of an operand: 128 rows x 128 B = 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

// copy a row-major [128][128] fp32 matrix into the swizzled K-blocked smem operand layout
__device__ void load_operand(uint8_t* dst, const float* __restrict__ src, int tid, int nthr) {
  for (int idx = tid; idx < M * (K / 4); idx += nthr) {       // 16-byte chunks
    const int r = idx / (K / 4), c = idx % (K / 4);
    const int kb = c / 8, cc = c % 8;
    const float4 v = *reinterpret_cast<const float4*>(src + r * K + c * 4);
    *reinterpret_cast<float4*>(dst + kb * TILE_BYTES + r * 128 + ((cc ^ (r & 7)) * 16)) = v;
  }
}

__global__ void __launch_bounds__(128) gemm_kernel(const float* A, const float* Alo, const float* B, const float* Blo,
                                                   float* D, int passes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                              // 4 x 16 KB
  uint8_t* sAlo = smem + 1 * NKB * TILE_BYTES;
  uint8_t* sB = smem + 2 * NKB * TILE_BYTES;
  uint8_t* sBlo = smem + 3 * NKB * TILE_BYTES;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  load_operand(sA, A, tid, 128);
  load_operand(sB, B, tid, 128);
  if (passes == 3) {
    load_operand(sAlo, Alo, tid, 128);
    load_operand(sBlo, Blo, tid, 128);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy smem writes -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  if (tid == 0) printf("tmem base 0x%08x  smem base 0x%08x (mod 1024 = %u)\n", tmem_d, smem_u32(smem), smem_u32(smem) & 1023u);
  {  // debug: pre-fill the accumulator with 7.0 so that "MMA never wrote" and "MMA wrote zeros" differ
    for (int c0 = 0; c0 < N; c0 += 1) {
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(__float_as_uint(7.0f)) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }

  // instruction descriptor: D=F32, A=B=TF32, K-major both, N=128, M=128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    uint32_t acc = 0;
    for (int pass = 0; pass < passes; ++pass) {
      const uint8_t* a = (pass == 1) ? sAlo : sA;      // pass 0: hi*hi, 1: lo*hi, 2: hi*lo
      const uint8_t* b = (pass == 2) ? sBlo : sB;
      for (int kb = 0; kb < NKB; ++kb) {
        const uint64_t da0 = make_desc(smem_u32(a + kb * TILE_BYTES));
        const uint64_t db0 = make_desc(smem_u32(b + kb * TILE_BYTES));
        for (int ks = 0; ks < KB / 8; ++ks) {          // UMMA_K = 8 tf32 = 32 bytes -> +2 in the address field
          mma_tf32_ss(tmem_d, da0 + 2 * ks, db0 + 2 * ks, idesc, acc);
          acc = 1;
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar))
                 : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar))
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows of D
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
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
    float* drow = D + (warp * 32 + lane) * N + c0;
#pragma unroll
    for (int j = 0; j < 32; ++j) drow[j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tmem_d));
}

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

int main() {
  std::vector<float> A(M * K), B(N * K), Alo(M * K), Blo(N * K), Ahi(M * K), Bhi(N * K), D(M * N);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX - 0.5f;
  for (auto& v : B) v = (float)rand() / RAND_MAX - 0.5f;
  for (int i = 0; i < M * K; ++i) { Ahi[i] = tf32_trunc(A[i]); Alo[i] = A[i] - Ahi[i]; }
  for (int i = 0; i < N * K; ++i) { Bhi[i] = tf32_trunc(B[i]); Blo[i] = B[i] - Bhi[i]; }
  std::vector<double> ref(M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
      ref[m * N + n] = s;
    }
  float *dA, *dAlo, *dB, *dBlo, *dD;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dAlo, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dBlo, N * K * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, Ahi.data(), M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dAlo, Alo.data(), M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bhi.data(), N * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dBlo, Blo.data(), N * K * 4, cudaMemcpyHostToDevice);
  const int smem = 4 * NKB * TILE_BYTES + 1024;
  cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int passes : {1, 3}) {
    cudaMemset(dD, 0, M * N * 4);
    gemm_kernel<<<1, 128, smem>>>(dA, dAlo, dB, dBlo, dD, passes);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(D[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("passes=%d: %s  max|D-ref| = %.3e  (max|ref| = %.3f, rel %.3e)  D[0]=%f ref[0]=%f\n", passes,
           cudaGetErrorString(e), maxerr, maxref, maxerr / maxref, D[0], ref[0]);
  }
  return 0;
}
