// Microbenchmark: issue rate of the complex-MAC inner loop in three register layouts, one warp
// per SMSP (the occupancy the chain kernels run at).  Prints cycles per real FMA per warp.
#include <cuda_runtime.h>
#include <cstdio>

constexpr int NCOL = 8;     // complex columns per thread (as DP=32, NQ=4)
constexpr int ITERS = 4096;

// (A) interleaved float2 (re,im) operands, scalar FFMA  -- what the kernels do now
__global__ void k_interleaved(const float2* __restrict__ mat, const float2* __restrict__ vec, float2* out, long long* cyc) {
  float2 A[NCOL], B[NCOL], x[NCOL];
  for (int c = 0; c < NCOL; ++c) { A[c] = mat[threadIdx.x * NCOL + c]; B[c] = mat[4096 + threadIdx.x * NCOL + c]; x[c] = vec[c]; }
  float2 a0 = {0, 0}, a1 = a0, b0 = a0, b1 = a0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < NCOL; c += 2) {
      a0.x = fmaf(A[c].x, x[c].x, a0.x); a0.y = fmaf(A[c].x, x[c].y, a0.y);
      a0.x = fmaf(-A[c].y, x[c].y, a0.x); a0.y = fmaf(A[c].y, x[c].x, a0.y);
      a1.x = fmaf(A[c + 1].x, x[c + 1].x, a1.x); a1.y = fmaf(A[c + 1].x, x[c + 1].y, a1.y);
      a1.x = fmaf(-A[c + 1].y, x[c + 1].y, a1.x); a1.y = fmaf(A[c + 1].y, x[c + 1].x, a1.y);
      b0.x = fmaf(B[c].x, x[c].x, b0.x); b0.y = fmaf(B[c].x, x[c].y, b0.y);
      b0.x = fmaf(-B[c].y, x[c].y, b0.x); b0.y = fmaf(B[c].y, x[c].x, b0.y);
      b1.x = fmaf(B[c + 1].x, x[c + 1].x, b1.x); b1.y = fmaf(B[c + 1].x, x[c + 1].y, b1.y);
      b1.x = fmaf(-B[c + 1].y, x[c + 1].y, b1.x); b1.y = fmaf(B[c + 1].y, x[c + 1].x, b1.y);
    }
    // rotate x so the loop is not hoisted
#pragma unroll
    for (int c = 0; c < NCOL; ++c) { x[c].x += 1e-9f * a0.x; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(a0.x + a1.x + b0.x + b1.x, a0.y + a1.y + b0.y + b1.y);
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// (B) packed FFMA2: matrix element (re,im) natural, x duplicated (xr,xr),(xi,xi); two packed accumulators
__global__ void k_packed(const float2* __restrict__ mat, const float2* __restrict__ vec, float2* out, long long* cyc) {
  float2 A[NCOL], B[NCOL], xr[NCOL], xi[NCOL];
  for (int c = 0; c < NCOL; ++c) { A[c] = mat[threadIdx.x * NCOL + c]; B[c] = mat[4096 + threadIdx.x * NCOL + c];
    xr[c] = make_float2(vec[c].x, vec[c].x); xi[c] = make_float2(vec[c].y, vec[c].y); }
  float2 a1 = {0, 0}, a2 = a1, b1 = a1, b2 = a1, a3 = a1, a4 = a1, b3 = a1, b4 = a1;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < NCOL; c += 2) {
      a1 = __ffma2_rn(A[c], xr[c], a1); a2 = __ffma2_rn(A[c], xi[c], a2);
      a3 = __ffma2_rn(A[c + 1], xr[c + 1], a3); a4 = __ffma2_rn(A[c + 1], xi[c + 1], a4);
      b1 = __ffma2_rn(B[c], xr[c], b1); b2 = __ffma2_rn(B[c], xi[c], b2);
      b3 = __ffma2_rn(B[c + 1], xr[c + 1], b3); b4 = __ffma2_rn(B[c + 1], xi[c + 1], b4);
    }
#pragma unroll
    for (int c = 0; c < NCOL; ++c) { xr[c].x += 1e-9f * a1.x; xr[c].y = xr[c].x; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] =
      make_float2(a1.x - a2.y + a3.x - a4.y + b1.x - b2.y + b3.x - b4.y, a1.y + a2.x + a3.y + a4.x + b1.y + b2.x + b3.y + b4.x);
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// (C) planar: separate re / im register arrays, scalar FFMA
__global__ void k_planar(const float2* __restrict__ mat, const float2* __restrict__ vec, float2* out, long long* cyc) {
  float Ar[NCOL], Ai[NCOL], Br[NCOL], Bi[NCOL], xr[NCOL], xi[NCOL];
  for (int c = 0; c < NCOL; ++c) { float2 a = mat[threadIdx.x * NCOL + c], bb = mat[4096 + threadIdx.x * NCOL + c];
    Ar[c] = a.x; Ai[c] = a.y; Br[c] = bb.x; Bi[c] = bb.y; xr[c] = vec[c].x; xi[c] = vec[c].y; }
  float ar0 = 0, ai0 = 0, ar1 = 0, ai1 = 0, br0 = 0, bi0 = 0, br1 = 0, bi1 = 0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < NCOL; c += 2) {
      ar0 = fmaf(Ar[c], xr[c], ar0); ai0 = fmaf(Ar[c], xi[c], ai0); ar0 = fmaf(-Ai[c], xi[c], ar0); ai0 = fmaf(Ai[c], xr[c], ai0);
      ar1 = fmaf(Ar[c+1], xr[c+1], ar1); ai1 = fmaf(Ar[c+1], xi[c+1], ai1); ar1 = fmaf(-Ai[c+1], xi[c+1], ar1); ai1 = fmaf(Ai[c+1], xr[c+1], ai1);
      br0 = fmaf(Br[c], xr[c], br0); bi0 = fmaf(Br[c], xi[c], bi0); br0 = fmaf(-Bi[c], xi[c], br0); bi0 = fmaf(Bi[c], xr[c], bi0);
      br1 = fmaf(Br[c+1], xr[c+1], br1); bi1 = fmaf(Br[c+1], xi[c+1], bi1); br1 = fmaf(-Bi[c+1], xi[c+1], br1); bi1 = fmaf(Bi[c+1], xr[c+1], bi1);
    }
#pragma unroll
    for (int c = 0; c < NCOL; ++c) { xr[c] += 1e-9f * ar0; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(ar0 + ar1 + br0 + br1, ai0 + ai1 + bi0 + bi1);
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// (D) shuffle + barrier latency: dependent chain of shfl.bfly ; and __syncthreads per iteration
__global__ void k_shfl(float* out, long long* cyc) {
  float v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(float* out, long long* cyc) {
  __shared__ float s[256];
  float v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) { s[threadIdx.x] = v; __syncthreads(); v += s[(threadIdx.x + 33) & 127]; }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float2 *mat, *vec, *out; long long* cyc; float* fo;
  cudaMalloc(&mat, 8192 * 8 * 2); cudaMalloc(&vec, 64 * 8); cudaMalloc(&out, 148 * 128 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&fo, 1024);
  cudaMemset(mat, 0, 8192 * 8 * 2); cudaMemset(vec, 0, 64 * 8);
  long long h;
  const double nf = (double)ITERS * NCOL * 2 * 4;  // real FMAs per thread
  for (int rep = 0; rep < 2; ++rep) {
    k_interleaved<<<148, 128>>>(mat, vec, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("interleaved float2 FFMA : %.3f cycles per FFMA (per warp)\n", h / nf);
    k_packed<<<148, 128>>>(mat, vec, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("packed FFMA2            : %.3f cycles per real FMA (%.3f per FFMA2)\n", h / nf, 2 * h / nf);
    k_planar<<<148, 128>>>(mat, vec, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("planar FFMA             : %.3f cycles per FFMA (per warp)\n", h / nf);
    k_shfl<<<148, 128>>>(fo, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("shfl.bfly+fadd dependent: %.1f cycles per (shfl+add)\n", (double)h / ITERS / 2);
    k_bar<<<148, 128>>>(fo, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS+bar.sync(128)+LDS+add: %.1f cycles per round\n", (double)h / ITERS);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
