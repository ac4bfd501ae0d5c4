// Microbenchmark of the "warp per clip" chain step (lane = matrix row, D = 32): per step every lane
// loads the whole state vector (uniform-address LDS), forms its row of L = N + s R, does the 32
// complex MACs, rotates by q, stores its element and __syncwarp()s.  Reports cycles per step for
// 1, 2, 4, 8 warps (clips) per SM.
#include <cuda_runtime.h>
#include <cstdio>

constexpr int D = 32;
constexpr int STEPS = 4096;

template <int LDSW>   // 128 or 64-bit loads
__global__ void k_chain(const float2* __restrict__ mats, const float* __restrict__ sv, float2* out, long long* cyc) {
  extern __shared__ __align__(16) float2 sm[];   // [warps][2][D]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float2* xs = sm + w * 2 * D;
  float2 Nr[D], Rr[D];
#pragma unroll
  for (int c = 0; c < D; ++c) { Nr[c] = mats[lane * D + c]; Rr[c] = mats[D * D + lane * D + c]; }
  xs[lane] = make_float2(1.0f / 6, 0.01f * lane);
  __syncwarp();
  int cur = 0;
  long long t0 = clock64();
  for (int k = 0; k < STEPS; ++k) {
    const float s = sv[k & 255];
    const float2* xv = xs + cur * D;
    float2 a0 = {0, 0}, a1 = {0, 0};
#pragma unroll
    for (int c = 0; c < D; c += 2) {
      float2 x0, x1;
      if (LDSW == 128) { const float4 v = *reinterpret_cast<const float4*>(xv + c); x0 = make_float2(v.x, v.y); x1 = make_float2(v.z, v.w); }
      else { x0 = xv[c]; x1 = xv[c + 1]; }
      const float2 l0 = make_float2(fmaf(s, Rr[c].x, Nr[c].x), fmaf(s, Rr[c].y, Nr[c].y));
      const float2 l1 = make_float2(fmaf(s, Rr[c + 1].x, Nr[c + 1].x), fmaf(s, Rr[c + 1].y, Nr[c + 1].y));
      a0.x = fmaf(l0.x, x0.x, a0.x); a0.y = fmaf(l0.x, x0.y, a0.y); a0.x = fmaf(-l0.y, x0.y, a0.x); a0.y = fmaf(l0.y, x0.x, a0.y);
      a1.x = fmaf(l1.x, x1.x, a1.x); a1.y = fmaf(l1.x, x1.y, a1.y); a1.x = fmaf(-l1.y, x1.y, a1.x); a1.y = fmaf(l1.y, x1.x, a1.y);
    }
    const float2 xp = make_float2(a0.x + a1.x, a0.y + a1.y);
    xs[(cur ^ 1) * D + lane] = make_float2(xp.x * 0.999f - xp.y * 0.01f, xp.x * 0.01f + xp.y * 0.999f);
    cur ^= 1;
    __syncwarp();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = xs[cur * D + lane];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float2 *mats, *out; float* sv; long long* cyc;
  cudaMalloc(&mats, 2 * D * D * 8); cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&sv, 1024); cudaMalloc(&cyc, 8);
  float2 h[2 * D * D];
  for (int i = 0; i < 2 * D * D; ++i) h[i] = make_float2((i % 33 == 0) ? 1.0f : 1e-3f * (i % 7), 1e-3f * (i % 5));
  cudaMemcpy(mats, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMemset(sv, 0, 1024);
  long long c;
  for (int warps : {1, 2, 4, 8}) {
    k_chain<128><<<148, warps * 32, warps * 2 * D * 8>>>(mats, sv, out, cyc);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.128 warps/SM=%d : %.1f cycles/step  (%.1f cycles per step-clip)\n", warps, (double)c / STEPS, (double)c / STEPS / warps);
    k_chain<64><<<148, warps * 32, warps * 2 * D * 8>>>(mats, sv, out, cyc);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.64  warps/SM=%d : %.1f cycles/step  (%.1f cycles per step-clip)\n", warps, (double)c / STEPS, (double)c / STEPS / warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
