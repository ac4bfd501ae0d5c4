// C ABI of libaudiomps.so (see include/audiomps.h).  Host-side launch logic only; the kernels
// are in amps_psi.cuh / amps_rho.cuh / amps_prep.cuh.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>

#include <dlfcn.h>

#include "../../include/audiomps.h"
#include "amps_prep.cuh"
#include "amps_psi.cuh"
#include "amps_psi_cluster.cuh"
#include "amps_rho.cuh"
#include "amps_psi_c4.cuh"
#include "amps_scan_tc.cuh"
#include "amps_tiles_tc.cuh"
#include "amps_sx_tc.cuh"

using namespace amps;

struct amps_ctx {
  int device = 0;
  int num_sms = 0;
  int c4_cap_fwd = 0, c4_cap_bwd = 0; // 4-CTA clusters of the D = 128 chain kernels resident at once (GPC granularity)
  bool use_clusters = true;   // AMPS_NO_CLUSTER=1 disables the 2-CTA cluster kernels
  char err[512] = {0};
  int64_t launches = 0;
  // host-entry buffers (amps_psi_loss_grad_host only; the device entry points own no memory)
  void* hbuf = nullptr;
  size_t hbuf_bytes = 0;
  cudaStream_t hstream = nullptr;
  // optional per-kernel timing (CUDA events on the launch stream)
  bool prof = false;
  cudaEvent_t ev[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  bool ev_valid[4] = {false, false, false, false};
  // data-parallel communicator (NCCL, resolved at run time; see amps_comm_init)
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  // checkpointed backward (amps_psi_loss_bwd_k): the forward replay of time window j-1 runs on this
  // second stream, next to the adjoint sweep of window j on the caller's stream
  cudaStream_t aux_stream = nullptr;
  cudaStream_t hi_stream = nullptr;   // highest priority: the latency-bound chain kernels of a partial wave (launch_waves)
  bool ckpt_overlap = true;   // AMPS_CKPT_SERIAL=1: replay on the caller's stream (measurement aid)
  int waves_mask = 3;         // AMPS_WAVES: bit 0 = forward, bit 1 = backward, bit 2 = D = 128 forward wave pipelining
  bool tc_tiles = true;       // AMPS_NO_TC_TILES=1: D = 33..64 gradient tiles inside the sequential kernel (FFMA)
  cudaEvent_t ev_fork = nullptr, ev_replay[2] = {nullptr, nullptr}, ev_bwd[2] = {nullptr, nullptr};
};

namespace {

int fail(amps_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CUDA_TRY(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(ctx, AMPS_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                   \
  } while (0)

#define LAUNCH_CHECK(ctx, name)                                                          \
  do {                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess)                                                               \
      return fail(ctx, AMPS_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    ctx->launches++;                                                                     \
  } while (0)

#define PROF_BEGIN(ctx, which, st)                                        \
  do {                                                                     \
    if ((ctx)->prof) cudaEventRecord((ctx)->ev[which][0], st);            \
  } while (0)
#define PROF_END(ctx, which, st)                                          \
  do {                                                                     \
    if ((ctx)->prof) {                                                     \
      cudaEventRecord((ctx)->ev[which][1], st);                           \
      (ctx)->ev_valid[which] = true;                                       \
    }                                                                      \
  } while (0)

// FP32 FMA issue-rate microbenchmark: 16 independent chains per thread
__global__ void fma_peak_kernel(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;  // never true; keeps the chains alive
}

// same with packed FFMA2 (fma.rn.f32x2): 8 independent packed chains per thread
__global__ void fma2_peak_kernel(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;
}

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// ---- NCCL, bound at run time ---------------------------------------------------------------
// The library does not link against NCCL: amps_comm_* resolve the few entry points they need from
// the libnccl.so.2 already loaded in the process (PyTorch ships one) or found by the dynamic loader
// (AMPS_NCCL_LIB overrides the name).  Types mirror nccl.h (2.x ABI).
struct NcclId {
  char internal[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* name = getenv("AMPS_NCCL_LIB");
  void* h = dlopen(name && name[0] ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) return nullptr;
  api.handle = h;
  return &api;
}
constexpr int kNcclFloat = 7, kNcclSum = 0;   // ncclFloat32, ncclSum

int padded_dim(int D) {
  if (D <= 0) return -1;
  if (D <= 8) return 8;
  if (D <= 16) return 16;
  if (D <= 32) return 32;
  if (D <= 64) return 64;
  if (D <= 128) return 128;   // row-split 4-CTA cluster kernels (amps_psi_c4.cuh): loss / gradient / trajectory
  return -1;
}
// time-chunk length (rescale period) of the kernels serving a padded bond dimension
int chunk_len_of(int DP) { return DP == 128 ? CH4 : CH; }

// A by value, or by device pointer when the caller supplies one (no host read-back per step)
AVal aval(const amps_params* p) { return AVal{p->A, p->A_dev}; }

// whole-clip launch of a scan kernel (no time window, no checkpoints)
SegFwd seg_full_f(int T) { return SegFwd{T, nullptr, 0, nullptr, 1, 0}; }
SegBwd seg_full_b(int T) { return SegBwd{T, nullptr, 0, 0}; }

template <int V>
using IC = std::integral_constant<int, V>;

// f(IC<DP>, IC<NQ>)
template <class F>
int dispatch_dp(int DP, F&& f) {
  switch (DP) {
    case 8: return f(IC<8>{}, IC<4>{});
    case 16: return f(IC<16>{}, IC<4>{});
    case 32: return f(IC<32>{}, IC<4>{});
    case 64: return f(IC<64>{}, IC<8>{});
    default: return AMPS_E_UNSUPPORTED;
  }
}

// time splits of the tensor-core tile kernel (amps_tiles_tc.cuh): enough CTAs for two waves, at least
// 1024 steps each; a function of the shape only (the workspace size must not depend on the context)
int tiles_nsplit(int DP, int B, int nsteps) {
  if (DP < 64 || B <= 0) return 1;
  int n = (2 * 148 + B - 1) / B;
  const int cap = nsteps / 1024;
  if (n > cap) n = cap;
  return n < 1 ? 1 : n;
}
int tiles_steps_per_split(int nsteps, int nsplit) {
  const int s = (nsteps + nsplit - 1) / nsplit;
  return ((s + TL_KS - 1) / TL_KS) * TL_KS;
}

// time splits of the expectation pass (amps_sx_tc.cuh): it keeps no per-split matrices, so it can use many more
// CTAs than the tile kernel -- about eight waves of one CTA per SM keep the tail wave under a tenth of the run
int sx_nsplit_of(int DP, int B, int nsteps) {
  if (DP < 64 || B <= 0) return 1;
  int n = (8 * 148 + B - 1) / B;
  const int cap = nsteps / 1024;
  if (n > cap) n = cap;
  return n < 1 ? 1 : n;
}
int sx_steps_per_split(int nsteps, int nsplit) {
  const int s = (nsteps + nsplit - 1) / nsplit;
  return ((s + 127) / 128) * 128;
}

struct PsiWs {
  size_t matN, matR, matRH, matS, spanel, psi0p, ttab, qtab, lossd;
  size_t traj, scales, G, gf, lam0, gAdir, Gtot, gftot, lam0tot, sptraj, ev, lossp;
  size_t total;
};

PsiWs psi_ws_layout(int DP, int B, int nsteps_tab, int T, bool save) {
  PsiWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  const size_t mat = (size_t)DP * DP * sizeof(float2);
  w.matN = take(mat);
  w.matR = take(mat);
  w.matRH = take(mat);
  w.matS = take(mat);
  w.spanel = take(DP == 128 ? 4 * mat : 0);   // S in real form, tf32 hi + lo, in the panel order of psi_sx2_tc_kernel
  w.psi0p = take((size_t)DP * sizeof(float2));
  w.ttab = take((size_t)(nsteps_tab > 0 ? nsteps_tab + 2 : 2) * sizeof(float));   // t_k, k = 0..nsteps_tab
  w.qtab = take((size_t)(nsteps_tab > 0 ? nsteps_tab : 1) * DP * sizeof(float2));
  w.lossd = take((size_t)(B > 0 ? B : 1) * sizeof(double));
  if (save) {
    const int nsteps = T - 1;
    const int chl = chunk_len_of(DP);
    const int nchunks = nsteps > 0 ? (nsteps + chl - 1) / chl : 0;
    w.traj = take((size_t)B * T * DP * sizeof(float2));
    w.scales = take((size_t)B * (nchunks > 0 ? nchunks : 1) * sizeof(float));
    w.G = take((size_t)B * tiles_nsplit(DP, B, nsteps) * 3 * mat);
    w.gf = take((size_t)B * DP * sizeof(float));
    w.lam0 = take((size_t)B * DP * sizeof(float2));
    w.gAdir = take((size_t)B * sizeof(double));
    w.Gtot = take(3 * mat);
    w.gftot = take((size_t)DP * sizeof(float));
    w.lam0tot = take((size_t)DP * sizeof(float2));
    {   // S x'_k and (E_k, |x_k|^2) from the forward
      w.sptraj = take((size_t)B * T * DP * sizeof(float2));   // S x'_k
      w.ev = take((size_t)B * T * sizeof(float2));            // (E_k, |x_k|^2)
      w.lossp = take((size_t)B * sx_nsplit_of(DP, B, nsteps) * sizeof(double));   // per-split loss sums (D > 32)
    }
  }
  w.total = off;
  return w;
}

int check_common(amps_ctx* ctx, const amps_params* p) {
  if (!ctx) return AMPS_E_INVALID;
  if (!p) return fail(ctx, AMPS_E_INVALID, "params is NULL");
  if (p->D <= 0) return fail(ctx, AMPS_E_INVALID, "bond dimension D=%d must be positive", p->D);
  if (!p->R_dev || !p->freqs_dev) return fail(ctx, AMPS_E_INVALID, "R_dev/freqs_dev is NULL");
  if (!p->A_dev && (!(p->A == p->A) || p->A == 0.0f)) return fail(ctx, AMPS_E_INVALID, "A must be non-zero");
  return AMPS_OK;
}

// mats + psi0 + q table into the workspace
int psi_prepare(amps_ctx* ctx, const amps_params* p, int DP, char* ws, const PsiWs& L,
                int nsteps_tab, cudaStream_t st) {
  const double cprime = -p->delta_t * (double)p->sigma * (double)p->sigma / 2.0;  // model.py:312
  // float32 time table of THIS call (model.py:16,281), into the caller's workspace: the backward and the
  // trajectory kernels read it from there, so no entry point depends on context state
  prep_ttab_kernel<<<1, 1024, 0, st>>>((float)p->delta_t, nsteps_tab + 1, (float*)(ws + L.ttab));
  LAUNCH_CHECK(ctx, "prep_ttab_kernel");
  prep_mats_kernel<<<(DP * DP + 255) / 256, 256, 0, st>>>(
      (const float2*)p->R_dev, p->D, DP, cprime, (float2*)(ws + L.matN), (float2*)(ws + L.matR),
      (float2*)(ws + L.matRH), (float2*)(ws + L.matS));
  LAUNCH_CHECK(ctx, "prep_mats_kernel");
  prep_pad_vec_kernel<<<1, DP, 0, st>>>((const float2*)p->psi0_dev, p->D, DP,
                                        (float2*)(ws + L.psi0p));
  LAUNCH_CHECK(ctx, "prep_pad_vec_kernel");
  if (nsteps_tab > 0) {
    const dim3 grid((nsteps_tab + QG - 1) / QG, (DP + QC - 1) / QC);
    prep_phase_tables_kernel<<<grid, 256, sizeof(PhaseSmem), st>>>(p->freqs_dev, p->D, DP, (const float*)(ws + L.ttab),
                                                                 nsteps_tab, (float2*)(ws + L.qtab), (float2*)nullptr);
    LAUNCH_CHECK(ctx, "prep_phase_tables_kernel");
  }
  return AMPS_OK;
}

// launch `kern` as `nclusters` thread-block clusters of CL CTAs
template <class K, class... Args>
cudaError_t launch_cluster(K kern, int nclusters, int CL, int threads, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * CL);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
constexpr int C4_CL = 4;
constexpr int C4_CHAIN_THREADS = 128;   // threads per CTA of the chain-only D = 128 kernels (launch_psi_fwd)

// plain launch, optionally with programmatic stream serialisation: the kernel may start once every CTA of the
// kernel queued before it in the stream has executed griddepcontrol.launch_dependents (or exited) -- it does NOT
// wait for that kernel, or the ones before it, to finish
template <class K, class... Args>
cudaError_t launch_pdl(K kern, dim3 grid, int threads, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// Rho tables of one call -- float32 t_k, q_k, and p_{k+1} when a lab-frame trajectory is wanted -- into
// the caller's workspace
int rho_tables(amps_ctx* ctx, const amps_params* p, int nsteps, bool need_p, char* ws, const RhoWs& L,
               cudaStream_t st) {
  prep_ttab_kernel<<<1, 1024, 0, st>>>((float)p->delta_t, nsteps + 1, (float*)(ws + L.ttab));
  LAUNCH_CHECK(ctx, "prep_ttab_kernel");
  if (nsteps > 0) {
    const dim3 grid((nsteps + QG - 1) / QG, (p->D + QC - 1) / QC);
    prep_phase_tables_kernel<<<grid, 256, sizeof(PhaseSmem), st>>>(p->freqs_dev, p->D, p->D, (const float*)(ws + L.ttab),
                                                                 nsteps, (float2*)(ws + L.qtab),
                                                                 need_p ? (float2*)(ws + L.ptab) : (float2*)nullptr);
    LAUNCH_CHECK(ctx, "prep_phase_tables_kernel");
  }
  return AMPS_OK;
}

}  // namespace

cudaError_t amps_set_all_func_attrs();   // defined with the launch helpers below
int amps_c4_cluster_capacity(bool bwd);  // idem
// second stream of a context (forward replay of the checkpointed backward, tensor-core pass of finished
// waves): lowest priority, so that the latency-bound chain kernels on the caller's stream get their SMs first
static cudaError_t create_aux_stream(cudaStream_t* s, bool highest = false) {
  int least = 0, greatest = 0;
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
  if (e != cudaSuccess) return e;
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, highest ? greatest : least);
}

// ------------------------------------------------------------------------------------------
extern "C" {

int amps_version(void) { return AMPS_VERSION; }

int amps_create(int device, amps_ctx** out) {
  if (!out) return AMPS_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return AMPS_E_CUDA;
  amps_ctx* ctx = new (std::nothrow) amps_ctx();
  if (!ctx) return AMPS_E_INVALID;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    delete ctx;
    return AMPS_E_CUDA;
  }
  const char* nc = getenv("AMPS_NO_CLUSTER");
  ctx->use_clusters = !(nc && nc[0] == '1');
  const char* cs = getenv("AMPS_CKPT_SERIAL");
  ctx->ckpt_overlap = !(cs && cs[0] == '1');
  if (const char* wm = getenv("AMPS_WAVES")) ctx->waves_mask = atoi(wm);
  const char* nt = getenv("AMPS_NO_TC_TILES");
  ctx->tc_tiles = !(nt && nt[0] == '1');
  // everything the device entry points need besides the caller's buffers is created HERE: kernel
  // attributes, the replay stream and its events (no allocation, no attribute call per launch)
  bool ok = amps_set_all_func_attrs() == cudaSuccess &&
            create_aux_stream(&ctx->aux_stream) == cudaSuccess &&
            create_aux_stream(&ctx->hi_stream, true) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i)
    ok = cudaEventCreateWithFlags(&ctx->ev_replay[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_bwd[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    amps_destroy(ctx);
    return AMPS_E_CUDA;
  }
  ctx->c4_cap_fwd = amps_c4_cluster_capacity(false);
  ctx->c4_cap_bwd = amps_c4_cluster_capacity(true);
  if (ctx->c4_cap_fwd <= 0) ctx->c4_cap_fwd = ctx->num_sms / 4;
  if (ctx->c4_cap_bwd <= 0) ctx->c4_cap_bwd = ctx->num_sms / 4;
  *out = ctx;
  return AMPS_OK;
}

int amps_destroy(amps_ctx* ctx) {
  if (!ctx) return AMPS_E_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->nccl_comm) amps_comm_destroy(ctx);
  if (ctx->hbuf) cudaFree(ctx->hbuf);
  if (ctx->hstream) cudaStreamDestroy(ctx->hstream);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->hi_stream) cudaStreamDestroy(ctx->hi_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_replay[i]) cudaEventDestroy(ctx->ev_replay[i]);
    if (ctx->ev_bwd[i]) cudaEventDestroy(ctx->ev_bwd[i]);
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j)
      if (ctx->ev[i][j]) cudaEventDestroy(ctx->ev[i][j]);
  delete ctx;
  return AMPS_OK;
}

int amps_set_profiling(amps_ctx* ctx, int enable) {
  if (!ctx) return AMPS_E_INVALID;
  if (enable && !ctx->ev[0][0]) {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 2; ++j) CUDA_TRY(ctx, cudaEventCreate(&ctx->ev[i][j]));
  }
  ctx->prof = enable != 0;
  return AMPS_OK;
}

int amps_get_kernel_ms(amps_ctx* ctx, int which, float* ms) {
  if (!ctx || !ms || which < 0 || which > 3) return AMPS_E_INVALID;
  if (!ctx->ev_valid[which]) return fail(ctx, AMPS_E_STATE, "no timed launch of kernel %d yet", which);
  CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev[which][1]));
  CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->ev[which][0], ctx->ev[which][1]));
  return AMPS_OK;
}

double amps_fma_peak_tflops(amps_ctx* ctx) { return amps_fma_peak_tflops2(ctx, 0); }

double amps_fma_peak_tflops2(amps_ctx* ctx, int packed) {
  if (!ctx) return -1.0;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.0;
  float* out = nullptr;
  if (cudaMalloc(&out, 4) != cudaSuccess) return -1.0;
  const int iters = 1 << 15, blocks = 148 * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, 0);
    if (packed)
      fma2_peak_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    else
      fma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ctx->launches++;
    const double flops = 2.0 * 16.0 * iters * (double)blocks * threads;
    if (ms > 0.f && rep > 0) best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
  }
  cudaFree(out);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

const char* amps_last_error(const amps_ctx* ctx) { return ctx ? ctx->err : "null context"; }
int64_t amps_launch_count(const amps_ctx* ctx) { return ctx ? ctx->launches : 0; }

// float32 running-sum time table on the HOST (same piecewise-exact generator as the device kernel)
int amps_time_table_host(double delta_t, int n, float* out) {
  if (n < 0 || (n > 0 && !out)) return AMPS_E_INVALID;
  TtSeg segs[TT_MAXSEG];
  const int nseg = tt_build((float)delta_t, n, out, segs, TT_MAXSEG);
  for (int s = 0; s < nseg; ++s)
    for (int j = 1; j <= segs[s].m; ++j)
      out[segs[s].k0 + j] = (float)((double)segs[s].t0 + (double)j * (double)segs[s].inc);
  return AMPS_OK;
}

size_t amps_psi_grad_count(int D) { return D > 0 ? (size_t)2 * D * D + 3 * (size_t)D + 2 : 0; }

size_t amps_psi_workspace_bytes(int D, int B, int T, int save_for_bwd) {
  const int DP = padded_dim(D);
  if (DP < 0 || B < 0 || T < 0) return 0;
  return psi_ws_layout(DP, B, T > 0 ? T - 1 : 0, T, save_for_bwd != 0).total;
}

}  // extern "C" (the scan launch helpers below are internal)

// ---- launch of the sequential scan kernels (whole clip or one time window) -----------------------
namespace {
struct FwdArgs {
  const float2 *matN, *matR, *matS, *qtab, *psi0p;
  const float* x;
  int T;                 // samples of the launch (window: steps + 1)
  AVal A;
  float* loss;
  double* lossd;
  float2* traj;          // null: nothing is kept for a backward
  float* scales;
  float2* sptraj;
  float2* ev;
  SegFwd seg;
  float4* spanel = nullptr;      // D = 128: panel-ordered tf32 split of S for the expectation pass
  bool build_panel = true;       // (built by the first wave only: later waves run next to a reader of it)
  bool paired = false;           // replay of the checkpointed backward (family_of)
  double* loss_part = nullptr;   // D > 32 saving forward: per-split loss sums of the expectation pass
  int sx_nsplit = 1, sx_sps = 0;
  bool allow_split = false;      // whole-batch call on the caller's stream: partial waves may be pipelined
};
struct BwdArgs {
  const float2 *matN, *matRH, *matS, *qtab;
  const float* ttab;
  const float* x;
  int T;
  AVal A;
  const float* w;
  const float2* traj;
  const float* scales;
  float2* G;
  float* gf;
  float2* lam0;
  double* gAdir;
  const float2* sptraj;
  const float2* ev;
  SegBwd seg;
  int tiles_nsplit = 1;        // partial tile sets per clip in G (D > 32: tensor-core tile kernel)
  int tiles_sps = 0;           // steps per split (multiple of 32)
  bool allow_split = false;
  bool paired = false;         // adjoint of the checkpointed backward, next to a replay (family_of)
};

// which kernel family serves (DP, B) on this context
enum class Fam { C4, CL, WS, UNI };
// paired: the launch is one of the two concurrent kernels of the checkpointed backward (replay of window j-1 next
// to the adjoint of window j).  Two 2-CTA-per-clip cluster kernels only run side by side while 4 B <= #SMs; above
// that (C1: 64 clips) the single-CTA family lets both fit (2 B <= #SMs) -- C1 at K = 2048: replay + adjoint
// 18.1 ms serialised on the cluster kernels, 13.9 ms overlapped on the single-CTA ones.
Fam family_of(const amps_ctx* ctx, int DP, int B, bool paired = false) {
  if (DP == 128) return Fam::C4;
  if (DP == 64) return Fam::UNI;
  if (paired && ctx->ckpt_overlap && 4 * B > ctx->num_sms) return Fam::WS;
  return (ctx->use_clusters && 2 * B <= ctx->num_sms) ? Fam::CL : Fam::WS;
}

// opt-in dynamic shared memory of every scan kernel, once per context (amps_create).  The carveout is
// pinned to "max shared" for all of them: an SM only co-hosts CTAs of two kernels (the checkpointed
// backward runs the replay next to the adjoint) if it does not have to be drained to change its
// L1/shared split.
template <class Kern>
cudaError_t set_smem(Kern k, size_t bytes, bool max_carveout = true) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess || !max_carveout || getenv("AMPS_DEFAULT_CARVEOUT")) return e;
  return cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
template <int DPc, int NQc>
cudaError_t set_attrs_ws() {
  cudaError_t e;
  if ((e = set_smem(psi_fwd_cl_kernel<DPc, NQc>, sizeof(FwdClSmem<DPc, NQc>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_cl_kernel<DPc, NQc>, bwd_cl_smem_bytes<DPc, NQc>())) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_kernel<DPc, NQc>, sizeof(FwdSmem<DPc, NQc>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_kernel<DPc, NQc>, sizeof(BwdSmem<DPc, NQc>))) != cudaSuccess) return e;
  return set_smem(psi_sample_kernel<DPc, NQc>, sizeof(SampleSmem<DPc>));
}
}  // namespace
// how many 4-CTA clusters of the D = 128 chain kernels the device holds at once: clusters do not span GPCs, so this
// is below num_sms / 4 per resident CTA (B200, 148 SMs: 32 clusters of the backward, one CTA per SM; 64 of the
// chain-only forward, two CTAs per SM)
int amps_c4_cluster_capacity(bool bwd) {
  using namespace amps;
  auto query = [&](auto kern, size_t smem) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(C4_CL * 64);
    cfg.blockDim = dim3(C4_CHAIN_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C4_CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    return n;
  };
  return bwd ? query(psi_bwd_c4_kernel<128, C4_CL, false, false, C4_CHAIN_THREADS>, sizeof(BwdC4Smem<128, C4_CL, true>))
             : query(psi_fwd_c4_kernel<128, C4_CL, false, true, true, C4_CHAIN_THREADS>, sizeof(FwdC4Smem<128, C4_CL>));
}
cudaError_t amps_set_all_func_attrs() {
  using namespace amps;
  cudaError_t e;
  if ((e = set_attrs_ws<8, 4>()) != cudaSuccess) return e;
  if ((e = set_attrs_ws<16, 4>()) != cudaSuccess) return e;
  if ((e = set_attrs_ws<32, 4>()) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_uni_kernel<64, 8>, sizeof(FwdSmemUni<64, 8>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_uni_kernel<64, 4>, sizeof(FwdSmemUni<64, 4>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_uni_kernel<64, 4, true>, sizeof(FwdSmemUni<64, 4>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_uni_kernel<64, 8>, sizeof(BwdSmemUni<64>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_uni_kernel<64, 8, false, false>, sizeof(BwdSmemUni<64, true>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_tiles_tc_kernel<64, 0>, sizeof(TilesSmem<64, 0>) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_tiles_tc_kernel<64, 1>, sizeof(TilesSmem<64, 1>) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_uni_kernel<64, 8, false, true>, sizeof(FwdSmemUni<64, 8, true>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_uni_kernel<64, 4, false, true>, sizeof(FwdSmemUni<64, 4, true>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_uni_kernel<64, 4, false, false>, sizeof(BwdSmemUni<64, true>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_sx_tc_kernel<64>, sizeof(SxSmem<64>) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_sx2_tc_kernel, sizeof(Sx2Smem) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_tiles_tc_kernel<128, 1>, sizeof(TilesSmem<128, 1>) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_tiles_tc_kernel<128, 2>, sizeof(TilesSmem<128, 2>) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_tiles_tc_kernel<128, 3>, sizeof(TilesSmem<128, 3>) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_sample_c4_kernel<128, C4_CL>, sizeof(SampleC4Smem<128, C4_CL>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_c4_kernel<128, C4_CL, false, false, C4_CHAIN_THREADS>, sizeof(BwdC4Smem<128, C4_CL, true>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_c4_kernel<128, C4_CL, false, true, true, C4_CHAIN_THREADS>, sizeof(FwdC4Smem<128, C4_CL>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_uni_kernel<64, 4, true, false>, sizeof(BwdSmemUni<64>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_uni_kernel<64, 4, true, true>, sizeof(BwdSmemUni<64>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_sample_kernel<64, 4>, sizeof(SampleSmem<64>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_fwd_c4_kernel<128, C4_CL, false>, sizeof(FwdC4Smem<128, C4_CL>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_bwd_c4_kernel<128, C4_CL, false>, sizeof(BwdC4Smem<128, C4_CL>))) != cudaSuccess) return e;
  if ((e = set_smem(psi_compose_tc_kernel<true>, sizeof(ScanTcSmem) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_compose_tc_kernel<false>, sizeof(ScanTcSmem) + 1024)) != cudaSuccess) return e;
  if ((e = set_smem(psi_scan_boundary_kernel, (2 * TC_N * TC_N + 3 * TC_N + 8) * sizeof(float))) != cudaSuccess) return e;
  if ((e = set_smem(psi_scan_boundary_bwd_kernel, (2 * TC_N * TC_N + TC_N) * sizeof(float))) != cudaSuccess) return e;
  if ((e = set_smem(prep_phase_tables_kernel, sizeof(PhaseSmem))) != cudaSuccess) return e;
  if ((e = set_smem(rho_bwd_kernel, rho_bwd_smem_bytes(RHO_MAX_D))) != cudaSuccess) return e;
  if ((e = set_smem(rho_scan_kernel<false>, rho_smem_bytes(RHO_MAX_D))) != cudaSuccess) return e;
  if ((e = set_smem(rho_scan_kernel<true>, rho_smem_bytes(RHO_MAX_D))) != cudaSuccess) return e;
  return cudaSuccess;
}
namespace {

// phase (tensor-core paths only): 0 = everything, 1 = the sequential chain kernel, 2 = the GEMM pass after it
int launch_psi_fwd(amps_ctx* ctx, int DP, int B, const FwdArgs& a, cudaStream_t st, int phase = 0) {
  const int nsteps = a.T - 1;
  const int chl = chunk_len_of(DP);
  const int nchunks = (nsteps + chl - 1) / chl;
  const Fam fam = family_of(ctx, DP, B, a.paired);
  auto sx_args = [&]() {
    SxArgs g{};
    g.matS = a.matS;
    g.sptraj = a.sptraj;
    g.ev = a.ev;
    g.x = a.x;
    g.loss_part = a.loss_part;
    g.spanel = a.spanel;
    g.T = a.T;
    g.xstride = a.seg.xstride;
    g.nsplit = a.sx_nsplit;
    g.steps_per_split = a.sx_sps;
    g.A = a.A;
    return g;
  };
  const bool sxo = ctx->tc_tiles && a.traj && a.sptraj && a.loss_part && (DP != 128 || a.spanel);   // chain-only forward + tensor-core expectation pass
  if (fam == Fam::C4 && sxo) {
    if (phase < 2) {
      if (a.build_panel) {
        psi_sx2_panel_kernel<<<8, SX_THREADS, 0, st>>>(a.matS, a.spanel);
        LAUNCH_CHECK(ctx, "psi_sx2_panel_kernel");
      }
      // chain-only kernels run 128 threads per CTA: FOUR lanes per row, 32 columns of N and R per thread (178 / 200
      // registers).  The step is a latency chain -- row-sum shuffles, the exchange between the CTAs, the waits of
      // every warp on it -- not a throughput problem: 512 threads (16 lanes, 4 shuffle levels, 16 warps per CTA)
      // 117.7 / 135.5 ms forward / backward chain at C3, 256 threads 95 / 121.6, 128 threads 87.3 / 111.5.
      CUDA_TRY(ctx, launch_cluster(psi_fwd_c4_kernel<128, C4_CL, false, true, true, C4_CHAIN_THREADS>, B, C4_CL,
                                   C4_CHAIN_THREADS, sizeof(FwdC4Smem<128, C4_CL>), st,
                                   a.matN, a.matR, a.matS, a.qtab, a.psi0p, a.x, a.T, a.A, a.loss, a.lossd, a.traj,
                                   a.scales, nchunks, (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg));
      LAUNCH_CHECK(ctx, "psi_fwd_c4_kernel<chain>");
    }
    if (phase == 1) return AMPS_OK;
    const SxArgs g = sx_args();
    if (phase != 4) {
      PROF_BEGIN(ctx, 2, st);
      CUDA_TRY(ctx, launch_pdl(psi_sx2_tc_kernel, dim3(B * g.nsplit), SX_BLOCK, sizeof(Sx2Smem) + 1024, st, phase == 3, g));
      PROF_END(ctx, 2, st);
      LAUNCH_CHECK(ctx, "psi_sx2_tc_kernel");
    }
    if (phase == 3) return AMPS_OK;
    psi_scan_sum_kernel<<<(B + 127) / 128, 128, 0, st>>>(a.loss_part, B, g.nsplit, a.loss, a.lossd);
    LAUNCH_CHECK(ctx, "psi_scan_sum_kernel");
    return AMPS_OK;
  }
  if (fam == Fam::C4) {   // rows of N, R, S split over a 4-CTA cluster per clip
    CUDA_TRY(ctx, launch_cluster(psi_fwd_c4_kernel<128, C4_CL, false>, B, C4_CL, 512, sizeof(FwdC4Smem<128, C4_CL>), st,
                                 a.matN, a.matR, a.matS, a.qtab, a.psi0p, a.x, a.T, a.A, a.loss, a.lossd, a.traj,
                                 a.scales, nchunks, (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg));
    LAUNCH_CHECK(ctx, "psi_fwd_c4_kernel");
    return AMPS_OK;
  }
  return dispatch_dp(DP, [&](auto dp, auto nq) -> int {
    constexpr int DPc = decltype(dp)::value, NQc = decltype(nq)::value;
    if constexpr (DPc <= 32) {   // warp-specialised chain/filler kernels
      if (fam == Fam::CL) {
        // enough idle SMs: one 2-CTA cluster per clip (chain CTA + filler CTA on a second SM)
        CUDA_TRY(ctx, launch_cluster(psi_fwd_cl_kernel<DPc, NQc>, B, 2, 2 * DPc * NQc + 32, sizeof(FwdClSmem<DPc, NQc>), st,
                                     a.matN, a.matR, a.matS, a.qtab, a.psi0p, a.x, a.T, a.A, a.loss, a.lossd,
                                     a.traj, a.scales, nchunks, a.sptraj, a.ev, a.seg));
        LAUNCH_CHECK(ctx, "psi_fwd_cl_kernel");
        return AMPS_OK;
      }
      psi_fwd_kernel<DPc, NQc><<<B, 2 * DPc * NQc, sizeof(FwdSmem<DPc, NQc>), st>>>(
          a.matN, a.matR, a.matS, a.qtab, a.psi0p, a.x, a.T, a.A, a.loss, a.lossd, a.traj, a.scales, nchunks,
          a.sptraj, a.ev, a.seg);
    } else if (sxo) {
      // chain-only forward (x'_k and |x_k|^2 stored), then S x'_k, E_k and the loss as ONE GEMM over the time
      // axis on the tensor cores, in place
      if (phase != 2) {
        // FOUR lanes per row (256 threads, 16 columns of N and R per thread, 104 registers, still two CTAs per SM):
        // one shuffle level less and an 8-warp instead of a 16-warp barrier on every step -- C4's chain
        // 39.6 -> 30.0 ms.  (The full kernel keeps 8 lanes per row: with the S slices it would need > 128 registers.)
        psi_fwd_uni_kernel<DPc, 4, false, true><<<B, DPc * 4, sizeof(FwdSmemUni<DPc, 4, true>), st>>>(
            a.matN, a.matR, a.matS, a.qtab, a.psi0p, a.x, a.T, a.A, a.loss, a.lossd, a.traj, a.scales, nchunks,
            (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg);
        LAUNCH_CHECK(ctx, "psi_fwd_uni_kernel<chain>");
      }
      if (phase == 1) return AMPS_OK;
      const SxArgs g = sx_args();
      PROF_BEGIN(ctx, 2, st);
      psi_sx_tc_kernel<DPc><<<B * g.nsplit, SX_BLOCK, sizeof(SxSmem<DPc>) + 1024, st>>>(g);
      PROF_END(ctx, 2, st);
      LAUNCH_CHECK(ctx, "psi_sx_tc_kernel");
      psi_scan_sum_kernel<<<(B + 127) / 128, 128, 0, st>>>(a.loss_part, B, g.nsplit, a.loss, a.lossd);
      LAUNCH_CHECK(ctx, "psi_scan_sum_kernel");
      return AMPS_OK;
    } else {
      // (the full forward -- loss evaluation, the checkpoint pass -- also with four lanes per row: 169 registers,
      // C4's checkpoint pass 69.4 -> 61.0 ms)
      psi_fwd_uni_kernel<DPc, 4><<<B, DPc * 4, sizeof(FwdSmemUni<DPc, 4>), st>>>(
          a.matN, a.matR, a.matS, a.qtab, a.psi0p, a.x, a.T, a.A, a.loss, a.lossd, a.traj, a.scales, nchunks,
          (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg);
    }
    LAUNCH_CHECK(ctx, "psi_fwd_kernel");
    return AMPS_OK;
  });
}

int launch_psi_bwd(amps_ctx* ctx, int DP, int B, const BwdArgs& a, cudaStream_t st, int phase = 0) {
  const int nsteps = a.T - 1;
  const int chl = chunk_len_of(DP);
  const int nchunks = (nsteps + chl - 1) / chl;
  const Fam fam = family_of(ctx, DP, B, a.paired);
  // the gradient tiles as GEMMs over the time axis on the tensor cores (after a chain-only adjoint sweep that
  // left mu_k in place of the consumed S x'_k)
  auto tiles_args = [&]() {
    TilesArgs g{};
    g.mu = a.sptraj;
    g.traj = a.traj;
    g.qtab = a.qtab;
    g.scales = a.scales;
    g.ev = a.ev;
    g.x = a.x;
    g.w = a.w;
    g.G = a.G;
    g.T = a.T;
    g.xstride = a.seg.xstride;
    g.nchunks = nchunks;
    g.chunk_len = chl;
    g.nsplit = a.tiles_nsplit;
    g.steps_per_split = a.tiles_sps;
    g.accumulate = a.seg.accumulate;
    g.A = a.A;
    return g;
  };
  if (fam == Fam::C4 && ctx->tc_tiles) {
    if (phase < 2) {
      CUDA_TRY(ctx, launch_cluster(psi_bwd_c4_kernel<128, C4_CL, false, false, C4_CHAIN_THREADS>, B, C4_CL,
                                   C4_CHAIN_THREADS, sizeof(BwdC4Smem<128, C4_CL, true>), st,
                                   a.matN, a.matRH, a.matS, a.qtab, a.ttab, a.x, a.T, a.A, a.w, a.traj, a.scales, nchunks,
                                   a.G, a.gf, a.lam0, a.gAdir, (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg,
                                   const_cast<float2*>(a.sptraj)));
      LAUNCH_CHECK(ctx, "psi_bwd_c4_kernel<chain>");
    }
    if (phase == 1) return AMPS_OK;
    const TilesArgs g = tiles_args();
    const dim3 grid(B * g.nsplit, 2);
    PROF_BEGIN(ctx, 3, st);
    if (phase != 4) {   // phase 3: G_N and G_R with programmatic serialisation (launch_waves); phase 4: G_E after them
      CUDA_TRY(ctx, launch_pdl(psi_tiles_tc_kernel<128, 2>, grid, TL_BLOCK, sizeof(TilesSmem<128, 2>) + 1024, st, phase == 3, g));
      CUDA_TRY(ctx, launch_pdl(psi_tiles_tc_kernel<128, 3>, grid, TL_BLOCK, sizeof(TilesSmem<128, 3>) + 1024, st, phase == 3, g));
    }
    if (phase != 3)
      CUDA_TRY(ctx, launch_pdl(psi_tiles_tc_kernel<128, 1>, grid, TL_BLOCK, sizeof(TilesSmem<128, 1>) + 1024, st, false, g));
    PROF_END(ctx, 3, st);
    LAUNCH_CHECK(ctx, "psi_tiles_tc_kernel<128,1>");
    return AMPS_OK;
  }
  if (fam == Fam::C4) {
    CUDA_TRY(ctx, launch_cluster(psi_bwd_c4_kernel<128, C4_CL, false>, B, C4_CL, 512, sizeof(BwdC4Smem<128, C4_CL>), st,
                                 a.matN, a.matRH, a.matS, a.qtab, a.ttab, a.x, a.T, a.A, a.w, a.traj, a.scales, nchunks,
                                 a.G, a.gf, a.lam0, a.gAdir, (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg,
                                 (float2*)nullptr));
    LAUNCH_CHECK(ctx, "psi_bwd_c4_kernel");
    return AMPS_OK;
  }
  return dispatch_dp(DP, [&](auto dp, auto nq) -> int {
    constexpr int DPc = decltype(dp)::value, NQc = decltype(nq)::value;
    if constexpr (DPc <= 32) {
      if (fam == Fam::CL) {
        CUDA_TRY(ctx, launch_cluster(psi_bwd_cl_kernel<DPc, NQc>, B, 2, 3 * DPc * NQc, bwd_cl_smem_bytes<DPc, NQc>(), st,
                                     a.matN, a.matRH, a.matS, a.qtab, a.ttab, a.x, a.T, a.A, a.w, a.traj, a.scales,
                                     nchunks, a.G, a.gf, a.lam0, a.gAdir, a.sptraj, a.ev, a.seg));
        LAUNCH_CHECK(ctx, "psi_bwd_cl_kernel");
        return AMPS_OK;
      }
      psi_bwd_kernel<DPc, NQc><<<B, 2 * DPc * NQc, sizeof(BwdSmem<DPc, NQc>), st>>>(
          a.matN, a.matRH, a.matS, a.qtab, a.ttab, a.x, a.T, a.A, a.w, a.traj, a.scales, nchunks, a.G, a.gf,
          a.lam0, a.gAdir, a.sptraj, a.ev, a.seg);
    } else if (ctx->tc_tiles) {
      // chain-only adjoint sweep (mu_k replaces the consumed S x'_k in place), then the gradient tiles as
      // GEMMs over the time axis on the tensor cores
      if (phase != 2) {
        // (four lanes per row, 256 threads, 118 registers -- as the chain-only forward: C4's chain 46.5 -> 40.5 ms)
        psi_bwd_uni_kernel<DPc, 4, false, false><<<B, DPc * 4, sizeof(BwdSmemUni<DPc, true>), st>>>(
            a.matN, a.matRH, a.matS, a.qtab, a.ttab, a.x, a.T, a.A, a.w, a.traj, a.scales, nchunks, a.G, a.gf,
            a.lam0, a.gAdir, (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg, const_cast<float2*>(a.sptraj));
        LAUNCH_CHECK(ctx, "psi_bwd_uni_kernel<chain>");
      }
      if (phase == 1) return AMPS_OK;
      const TilesArgs g = tiles_args();
      PROF_BEGIN(ctx, 3, st);
      psi_tiles_tc_kernel<DPc, 0><<<B * g.nsplit, TL_BLOCK, sizeof(TilesSmem<DPc, 0>) + 1024, st>>>(g);
      LAUNCH_CHECK(ctx, "psi_tiles_tc_kernel<0>");
      psi_tiles_tc_kernel<DPc, 1><<<B * g.nsplit, TL_BLOCK, sizeof(TilesSmem<DPc, 1>) + 1024, st>>>(g);
      PROF_END(ctx, 3, st);
      LAUNCH_CHECK(ctx, "psi_tiles_tc_kernel<1>");
      return AMPS_OK;
    } else {
      psi_bwd_uni_kernel<DPc, NQc><<<B, DPc * NQc, sizeof(BwdSmemUni<DPc>), st>>>(
          a.matN, a.matRH, a.matS, a.qtab, a.ttab, a.x, a.T, a.A, a.w, a.traj, a.scales, nchunks, a.G, a.gf,
          a.lam0, a.gAdir, (const float2*)nullptr, 0, 0, a.sptraj, a.ev, a.seg);
    }
    LAUNCH_CHECK(ctx, "psi_bwd_kernel");
    return AMPS_OK;
  });
}

// Batches beyond one wave of chain CTAs (D = 64: one CTA per SM in the backward, two in the forward; C4: 256 clips =
// 148 + 108): the last, partial wave leaves SMs idle.  The batch is cut into the full waves (A) and the remainder
// (R): chain(A); then chain(R) next to gemm(A) -- the tensor-core pass of the clips already done, on the caller's
// stream, in the SMs chain(R) (highest-priority stream) leaves free; then gemm(R).  Results are identical (every
// kernel works per clip).  Measured: C4's backward -3.5 ms.  D = 128: see inside.
template <class Args, class Shift, class Launch>
int launch_waves(amps_ctx* ctx, int DP, int B, const Args& a, cudaStream_t st, bool tensor_path, int cap, Shift shift,
                 Launch launch) {
  if (!tensor_path || !a.allow_split || ctx->prof || !ctx->ckpt_overlap || DP < 64 || cap <= 0 || B <= cap ||
      st == ctx->aux_stream)
    return launch(B, a, st, 0);
  int rc;
  if (DP == 128) {
    // D = 128: the device holds `cap` 4-CTA clusters (B200: 32 = 128 SMs; clusters do not span GPCs), so EVERY wave
    // leaves SMs idle.  All in the caller's stream: chain(wave w), then the GEMM pass of wave w-1 launched with
    // programmatic stream serialisation -- it starts when every CTA of chain(w) has executed
    // griddepcontrol.launch_dependents (its first instruction), i.e. once the clusters are placed, and works in the
    // SMs they leave free.  (Two streams lose here: the GEMM CTAs win the race for the SMs and the clusters of
    // the next wave then wait for four free SMs of one GPC -- C3 313 -> 352 ms.)  The backward overlaps only
    // G_N and G_R (two thirds of the tile work: all three would outlast a wave on the spare SMs); G_E follows.
    const int nw = (B + cap - 1) / cap;
    for (int w = 0; w < nw; ++w) {
      const int b0 = w * cap, n = B - b0 < cap ? B - b0 : cap;
      if ((rc = launch(n, shift(a, b0), st, 1))) return rc;
      if (w > 0 && (rc = launch(cap, shift(a, b0 - cap), st, 3))) return rc;
    }
    const int bl = (nw - 1) * cap;
    if ((rc = launch(B - bl, shift(a, bl), st, 2))) return rc;
    for (int w = 0; w + 1 < nw; ++w)
      if ((rc = launch(cap, shift(a, w * cap), st, 4))) return rc;
    return AMPS_OK;
  }
  const int R = B % cap;
  if (R == 0) return launch(B, a, st, 0);
  const int Afull = B - R;
  const Args aR = shift(a, Afull);
  // (the default priority of a stream is the LOWEST: the remainder's chain kernel gets the context's
  // highest-priority stream so that its CTAs / clusters are placed before the GEMM CTAs of the finished waves)
  cudaStream_t sh = ctx->hi_stream;
  if ((rc = launch(Afull, a, st, 1))) return rc;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
  CUDA_TRY(ctx, cudaStreamWaitEvent(sh, ctx->ev_fork, 0));
  if ((rc = launch(R, aR, sh, 1))) return rc;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_replay[0], sh));
  if ((rc = launch(Afull, a, st, 2))) return rc;
  CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_replay[0], 0));
  if ((rc = launch(R, aR, st, 2))) return rc;
  return AMPS_OK;
}
int launch_psi_fwd_waves(amps_ctx* ctx, int DP, int B, const FwdArgs& a, cudaStream_t st) {
  const bool tensor_path = ctx->tc_tiles && a.traj && a.sptraj && a.loss_part && !a.seg.x0 && !a.seg.ckpt;
  const int chl = chunk_len_of(DP), nchunks = (a.T - 1 + chl - 1) / chl;
  auto shift = [&](const FwdArgs& f, int b0) {
    FwdArgs r = f;
    r.x += (size_t)b0 * f.seg.xstride;
    r.loss += b0;
    if (r.lossd) r.lossd += b0;
    r.traj += (size_t)b0 * f.T * DP;
    r.scales += (size_t)b0 * nchunks;
    r.sptraj += (size_t)b0 * f.T * DP;
    r.ev += (size_t)b0 * f.T;
    r.loss_part += (size_t)b0 * f.sx_nsplit;
    r.build_panel = b0 == 0;
    return r;
  };
  // the chain-only D = 64 forward fits two CTAs per SM (64 registers, 85 KB)
  // (D = 128: the chain-only forward runs two CTAs per SM, its clusters cover every SM and the GEMM CTAs -- one per
  // SM, 165 KB -- find no room next to them: pipelined 134 ms, plain 129 ms.  AMPS_WAVES=7 forces it.)
  const bool pipe = DP == 128 ? (ctx->waves_mask & 4) != 0 : (ctx->waves_mask & 1) != 0;
  return launch_waves(ctx, DP, B, a, st, tensor_path && pipe, DP == 128 ? ctx->c4_cap_fwd : 2 * ctx->num_sms, shift,
                      [&](int Bp, const FwdArgs& ap, cudaStream_t s, int ph) { return launch_psi_fwd(ctx, DP, Bp, ap, s, ph); });
}
int launch_psi_bwd_waves(amps_ctx* ctx, int DP, int B, const BwdArgs& a, cudaStream_t st) {
  const bool tensor_path = ctx->tc_tiles && !a.seg.lam_end && !a.seg.accumulate;
  const int chl = chunk_len_of(DP), nchunks = (a.T - 1 + chl - 1) / chl;
  auto shift = [&](const BwdArgs& f, int b0) {
    BwdArgs r = f;
    r.x += (size_t)b0 * f.seg.xstride;
    r.w += b0;
    r.traj += (size_t)b0 * f.T * DP;
    r.scales += (size_t)b0 * nchunks;
    r.G += (size_t)b0 * f.tiles_nsplit * 3 * DP * DP;
    r.gf += (size_t)b0 * DP;
    r.lam0 += (size_t)b0 * DP;
    r.gAdir += b0;
    r.sptraj += (size_t)b0 * f.T * DP;
    r.ev += (size_t)b0 * f.T;
    return r;
  };
  // (D = 64: the chain-only backward fits two CTAs per SM, like the forward)
  return launch_waves(ctx, DP, B, a, st, tensor_path && (ctx->waves_mask & 2), DP == 128 ? ctx->c4_cap_bwd : 2 * ctx->num_sms, shift,
                      [&](int Bp, const BwdArgs& ap, cudaStream_t s, int ph) { return launch_psi_bwd(ctx, DP, Bp, ap, s, ph); });
}

// clip reduction + finalize: packed effective-parameter gradient
int psi_finalize(amps_ctx* ctx, const amps_params* p, int DP, int B, char* ws, const PsiWs& L, const float* w_dev,
                 float* grad_dev, cudaStream_t st, int g_parts = 0) {
  const double cprime = -p->delta_t * (double)p->sigma * (double)p->sigma / 2.0;
  const int total = 3 * DP * DP + 2 * DP;
  psi_reduce_clips_kernel<<<(total + 127) / 128, 128, 0, st>>>(
      (const float2*)(ws + L.G), (const float*)(ws + L.gf), (const float2*)(ws + L.lam0), B, DP,
      (float2*)(ws + L.Gtot), (float*)(ws + L.gftot), (float2*)(ws + L.lam0tot), 1, g_parts);
  LAUNCH_CHECK(ctx, "psi_reduce_clips_kernel");
  psi_grad_finalize_kernel<<<p->D, 128, 0, st>>>(
      (const float2*)(ws + L.Gtot), (const float*)(ws + L.gftot), (const float2*)(ws + L.lam0tot),
      (const double*)(ws + L.gAdir), (const double*)(ws + L.lossd), w_dev, B,
      (const float2*)(ws + L.matR), p->D, DP, cprime, aval(p), grad_dev);
  LAUNCH_CHECK(ctx, "psi_grad_finalize_kernel");
  return AMPS_OK;
}

int check_fwd_args(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T, const float* loss_dev,
                   const void* ws_dev) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (!p->psi0_dev) return fail(ctx, AMPS_E_INVALID, "psi0_dev is NULL");
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d (need B>=0, T>=1)", B, T);
  if (B > 0 && (!x_dev || !loss_dev || !ws_dev)) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  if (padded_dim(p->D) < 0) return fail(ctx, AMPS_E_UNSUPPORTED, "bond dimension %d > 128 not supported", p->D);
  return AMPS_OK;
}
}  // namespace

extern "C" {

int amps_psi_loss_fwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      float* loss_dev, void* ws_dev, size_t ws_bytes, int save_for_bwd,
                      void* stream) {
  int rc = check_fwd_args(ctx, p, x_dev, B, T, loss_dev, ws_dev);
  if (rc) return rc;
  if (B == 0) return AMPS_OK;
  const int DP = padded_dim(p->D);
  const bool save = save_for_bwd != 0;
  const PsiWs L = psi_ws_layout(DP, B, T - 1, T, save);
  if (ws_bytes < L.total)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)ws_dev;
  rc = psi_prepare(ctx, p, DP, ws, L, T - 1, st);
  if (rc) return rc;
  FwdArgs a{(const float2*)(ws + L.matN), (const float2*)(ws + L.matR), (const float2*)(ws + L.matS),
            (const float2*)(ws + L.qtab), (const float2*)(ws + L.psi0p), x_dev, T, aval(p), loss_dev,
            (double*)(ws + L.lossd), save ? (float2*)(ws + L.traj) : nullptr,
            save ? (float*)(ws + L.scales) : nullptr, save ? (float2*)(ws + L.sptraj) : nullptr,
            save ? (float2*)(ws + L.ev) : nullptr, seg_full_f(T)};
  if (save) {
    a.loss_part = (double*)(ws + L.lossp);
    a.spanel = (float4*)(ws + L.spanel);
    a.sx_nsplit = sx_nsplit_of(DP, B, T - 1);
    a.sx_sps = sx_steps_per_split(T - 1, a.sx_nsplit);
    a.allow_split = true;
  }
  PROF_BEGIN(ctx, 0, st);
  rc = launch_psi_fwd_waves(ctx, DP, B, a, st);
  PROF_END(ctx, 0, st);
  return rc;
}

int amps_psi_loss_bwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev,
                      void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  if (!grad_dev) return fail(ctx, AMPS_E_INVALID, "grad_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ng = amps_psi_grad_count(p->D);
  if (B == 0) {
    CUDA_TRY(ctx, cudaMemsetAsync(grad_dev, 0, ng * sizeof(float), st));
    return AMPS_OK;
  }
  if (!x_dev || !w_dev || !ws_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  const int DP = padded_dim(p->D);
  if (DP < 0) return fail(ctx, AMPS_E_UNSUPPORTED, "bond dimension %d > 128 not supported", p->D);
  const PsiWs L = psi_ws_layout(DP, B, T - 1, T, true);
  if (ws_bytes < L.total)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  char* ws = (char*)ws_dev;
  BwdArgs a{(const float2*)(ws + L.matN), (const float2*)(ws + L.matRH), (const float2*)(ws + L.matS),
            (const float2*)(ws + L.qtab), (const float*)(ws + L.ttab), x_dev, T, aval(p), w_dev,
            (const float2*)(ws + L.traj), (const float*)(ws + L.scales), (float2*)(ws + L.G), (float*)(ws + L.gf),
            (float2*)(ws + L.lam0), (double*)(ws + L.gAdir), (const float2*)(ws + L.sptraj),
            (const float2*)(ws + L.ev), seg_full_b(T)};
  a.tiles_nsplit = tiles_nsplit(DP, B, T - 1);
  a.tiles_sps = tiles_steps_per_split(T - 1, a.tiles_nsplit);
  a.allow_split = true;
  const bool tc = DP >= 64 && ctx->tc_tiles;
  PROF_BEGIN(ctx, 1, st);
  rc = launch_psi_bwd_waves(ctx, DP, B, a, st);
  PROF_END(ctx, 1, st);
  if (rc) return rc;
  return psi_finalize(ctx, p, DP, B, ws, L, w_dev, grad_dev, st, tc ? B * a.tiles_nsplit : 0);
}

}  // extern "C"

// ---- checkpointed loss / gradient: the state is kept every K steps only ---------------------------
// Forward: one whole-clip launch that stores x at the start of every K-step window (nothing else).
// Backward, windows j = last .. 0: the forward kernel REPLAYS window j from its checkpoint into one of two
// window-sized trajectory buffers (x_k, S x'_k, (E_k,|x_k|^2), c_k), the adjoint kernel sweeps the window
// from the adjoint the next window left (Lam of its start state) and continues the gradient sums.  The
// replay of window j-1 runs on the context's second stream while the adjoint of window j runs on the
// caller's: where the batch leaves room on the SMs (C1: a 160-thread replay CTA next to a 384-thread
// adjoint CTA) the recompute hides behind the latency-bound adjoint chain.
namespace {
struct CkWs {
  PsiWs base;               // mats, psi0p, ttab, qtab, lossd | G, gf, lam0, gAdir, Gtot, gftot, lam0tot
  size_t ckpt;              // c64 [B][nwin][DP]
  size_t loss_scr, lossd_scr;
  size_t traj[2], sptraj[2], ev[2], scales[2];
  size_t total;
  int W, nwin;
};
// effective checkpoint interval: K rounded up to whole rescale chunks of the kernels serving D
int ck_interval(int DP, int K) {
  const int chl = chunk_len_of(DP);
  return ((K + chl - 1) / chl) * chl;
}
CkWs ck_ws_layout(int DP, int B, int T, int K) {
  CkWs w{};
  const int nsteps = T - 1;
  w.W = ck_interval(DP, K);
  w.nwin = nsteps > 0 ? (nsteps + w.W - 1) / w.W : 1;
  w.base = psi_ws_layout(DP, B, nsteps, T, false);
  size_t off = w.base.total;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  const size_t mat = (size_t)DP * DP * sizeof(float2);
  const int Wt = (nsteps < w.W ? nsteps : w.W) + 1;   // samples of a window
  const int wch = (w.W + chunk_len_of(DP) - 1) / chunk_len_of(DP);
  w.ckpt = take((size_t)B * w.nwin * DP * sizeof(float2));
  w.loss_scr = take((size_t)B * sizeof(float));
  w.lossd_scr = take((size_t)B * sizeof(double));
  w.base.lossp = take((size_t)B * sx_nsplit_of(DP, B, w.W) * sizeof(double));
  for (int i = 0; i < 2; ++i) {
    w.traj[i] = take((size_t)B * Wt * DP * sizeof(float2));
    w.sptraj[i] = take((size_t)B * Wt * DP * sizeof(float2));
    w.ev[i] = take((size_t)B * Wt * sizeof(float2));
    w.scales[i] = take((size_t)B * wch * sizeof(float));
  }
  w.base.G = take((size_t)B * tiles_nsplit(DP, B, w.W) * 3 * mat);
  w.base.gf = take((size_t)B * DP * sizeof(float));
  w.base.lam0 = take((size_t)B * DP * sizeof(float2));
  w.base.gAdir = take((size_t)B * sizeof(double));
  w.base.Gtot = take(3 * mat);
  w.base.gftot = take((size_t)DP * sizeof(float));
  w.base.lam0tot = take((size_t)DP * sizeof(float2));
  w.total = off;
  return w;
}
}  // namespace

extern "C" {

int amps_psi_ckpt_interval(int D, int K) {
  const int DP = padded_dim(D);
  if (DP < 0 || K < 1) return 0;
  return K == 1 ? 1 : ck_interval(DP, K);
}

size_t amps_psi_workspace_bytes_k(int D, int B, int T, int K) {
  const int DP = padded_dim(D);
  if (DP < 0 || B < 0 || T < 0 || K < 1) return 0;
  if (K == 1) return psi_ws_layout(DP, B, T > 0 ? T - 1 : 0, T, true).total;
  return ck_ws_layout(DP, B, T > 0 ? T : 1, K).total;
}

int amps_psi_loss_fwd_k(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T, int K,
                        float* loss_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (K == 1) return amps_psi_loss_fwd(ctx, p, x_dev, B, T, loss_dev, ws_dev, ws_bytes, 1, stream);
  int rc = check_fwd_args(ctx, p, x_dev, B, T, loss_dev, ws_dev);
  if (rc) return rc;
  if (K < 1) return fail(ctx, AMPS_E_INVALID, "checkpoint interval K=%d must be >= 1", K);
  if (B == 0) return AMPS_OK;
  const int DP = padded_dim(p->D);
  const CkWs L = ck_ws_layout(DP, B, T, K);
  if (ws_bytes < L.total)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)ws_dev;
  rc = psi_prepare(ctx, p, DP, ws, L.base, T - 1, st);
  if (rc) return rc;
  SegFwd seg = seg_full_f(T);
  seg.ckpt = (float2*)(ws + L.ckpt);
  seg.ck_chunks = L.W / chunk_len_of(DP);
  seg.ck_stride = L.nwin * DP;
  FwdArgs a{(const float2*)(ws + L.base.matN), (const float2*)(ws + L.base.matR), (const float2*)(ws + L.base.matS),
            (const float2*)(ws + L.base.qtab), (const float2*)(ws + L.base.psi0p), x_dev, T, aval(p), loss_dev,
            (double*)(ws + L.base.lossd), nullptr, nullptr, nullptr, nullptr, seg};
  PROF_BEGIN(ctx, 0, st);
  rc = launch_psi_fwd(ctx, DP, B, a, st);
  PROF_END(ctx, 0, st);
  return rc;
}

int amps_psi_loss_bwd_k(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T, int K,
                        const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev, void* stream) {
  if (K == 1) return amps_psi_loss_bwd(ctx, p, x_dev, B, T, w_dev, ws_dev, ws_bytes, grad_dev, stream);
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  if (K < 1) return fail(ctx, AMPS_E_INVALID, "checkpoint interval K=%d must be >= 1", K);
  if (!grad_dev) return fail(ctx, AMPS_E_INVALID, "grad_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    CUDA_TRY(ctx, cudaMemsetAsync(grad_dev, 0, amps_psi_grad_count(p->D) * sizeof(float), st));
    return AMPS_OK;
  }
  if (!x_dev || !w_dev || !ws_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  const int DP = padded_dim(p->D);
  if (DP < 0) return fail(ctx, AMPS_E_UNSUPPORTED, "bond dimension %d > 128 not supported", p->D);
  const CkWs L = ck_ws_layout(DP, B, T, K);
  if (ws_bytes < L.total)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  char* ws = (char*)ws_dev;
  const int nsteps = T - 1;
  if (nsteps == 0) {   // nothing to sweep: Lam_0 = 0, all sums zero
    const size_t mat = (size_t)DP * DP * sizeof(float2);
    CUDA_TRY(ctx, cudaMemsetAsync(ws + L.base.G, 0, (size_t)B * tiles_nsplit(DP, B, L.W) * 3 * mat, st));
    CUDA_TRY(ctx, cudaMemsetAsync(ws + L.base.gf, 0, (size_t)B * DP * sizeof(float), st));
    CUDA_TRY(ctx, cudaMemsetAsync(ws + L.base.lam0, 0, (size_t)B * DP * sizeof(float2), st));
    CUDA_TRY(ctx, cudaMemsetAsync(ws + L.base.gAdir, 0, (size_t)B * sizeof(double), st));
    return psi_finalize(ctx, p, DP, B, ws, L.base, w_dev, grad_dev, st);
  }
  cudaStream_t s2 = ctx->ckpt_overlap ? ctx->aux_stream : st;
  const int ck_nsplit = tiles_nsplit(DP, B, L.W), ck_sps = tiles_steps_per_split(L.W, ck_nsplit);
  const bool tc = DP >= 64 && ctx->tc_tiles;
  const float2* qtab = (const float2*)(ws + L.base.qtab);
  const float* ttab = (const float*)(ws + L.base.ttab);
  auto wlen = [&](int j) { return (nsteps - j * L.W < L.W) ? nsteps - j * L.W : L.W; };
  auto replay = [&](int j) -> int {     // forward of window j from its checkpoint, on the second stream
    const int i = j & 1;
    SegFwd seg{T, (const float2*)(ws + L.ckpt) + (size_t)j * DP, L.nwin * DP, nullptr, 1, 0};
    FwdArgs a{(const float2*)(ws + L.base.matN), (const float2*)(ws + L.base.matR), (const float2*)(ws + L.base.matS),
              qtab + (size_t)j * L.W * DP, (const float2*)(ws + L.base.psi0p), x_dev + (size_t)j * L.W, wlen(j) + 1,
              aval(p), (float*)(ws + L.loss_scr), (double*)(ws + L.lossd_scr), (float2*)(ws + L.traj[i]),
              (float*)(ws + L.scales[i]), (float2*)(ws + L.sptraj[i]), (float2*)(ws + L.ev[i]), seg};
    a.loss_part = (double*)(ws + L.base.lossp);
    a.spanel = (float4*)(ws + L.base.spanel);
    a.sx_nsplit = sx_nsplit_of(DP, B, L.W);
    a.sx_sps = sx_steps_per_split(L.W, a.sx_nsplit);
    a.paired = true;
    return launch_psi_fwd(ctx, DP, B, a, s2);
  };
  auto adjoint = [&](int j) -> int {
    const int i = j & 1;
    const bool last = j == L.nwin - 1;
    SegBwd seg{T, last ? nullptr : (const float2*)(ws + L.base.lam0), last ? 0 : 1, j > 0 ? 1 : 0};
    BwdArgs a{(const float2*)(ws + L.base.matN), (const float2*)(ws + L.base.matRH), (const float2*)(ws + L.base.matS),
              qtab + (size_t)j * L.W * DP, ttab + (size_t)j * L.W, x_dev + (size_t)j * L.W, wlen(j) + 1, aval(p), w_dev,
              (const float2*)(ws + L.traj[i]), (const float*)(ws + L.scales[i]), (float2*)(ws + L.base.G),
              (float*)(ws + L.base.gf), (float2*)(ws + L.base.lam0), (double*)(ws + L.base.gAdir),
              (const float2*)(ws + L.sptraj[i]), (const float2*)(ws + L.ev[i]), seg};
    a.tiles_nsplit = ck_nsplit;     // fixed by the window length W: every window fills the same partial slots
    a.tiles_sps = ck_sps;
    a.paired = true;
    return launch_psi_bwd(ctx, DP, B, a, st);
  };
  PROF_BEGIN(ctx, 1, st);
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));          // the replay stream joins behind the forward
  CUDA_TRY(ctx, cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
  if ((rc = replay(L.nwin - 1))) return rc;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_replay[(L.nwin - 1) & 1], s2));
  for (int j = L.nwin - 1; j >= 0; --j) {
    CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_replay[j & 1], 0));
    if (j > 0) {
      // window j-1 lands in the buffer the adjoint of window j+1 has just read
      if (j + 1 < L.nwin) CUDA_TRY(ctx, cudaStreamWaitEvent(s2, ctx->ev_bwd[(j + 1) & 1], 0));
      if ((rc = replay(j - 1))) return rc;
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_replay[(j - 1) & 1], s2));
    }
    if ((rc = adjoint(j))) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_bwd[j & 1], st));
  }
  PROF_END(ctx, 1, st);
  return psi_finalize(ctx, p, DP, B, ws, L.base, w_dev, grad_dev, st, tc ? B * ck_nsplit : 0);
}

}  // extern "C"

// ---- parallel-in-time loss / gradient (tcgen05 operator scan), D <= 64 -------------------------
namespace {
constexpr int kScanCtas = 148;   // virtual clips aim at one CTA per B200 SM (fixed: the workspace
                                 // size must not depend on the context)
struct ScanWs {
  PsiWs base;
  size_t ops, ystart, lossv, rnv;
  size_t traj, sptraj, ev, scales, G, gf, lam0, gAdir, lamend, wv, Gtot, gftot, lam0tot;
  size_t total;
  int nvc, m_steps;
};
ScanWs scan_ws_layout(int B, int T, bool save) {
  ScanWs w{};
  w.base = psi_ws_layout(64, B, T - 1, T, false);
  const int nsteps = T - 1;
  int nvc = B > 0 && B < kScanCtas ? kScanCtas / B : 1;       // virtual clips per clip: at most ONE wave of CTAs
  int m = nsteps > 0 ? (nsteps + nvc - 1) / nvc : CH;
  m = ((m + CH - 1) / CH) * CH;                                // whole 32-step chunks
  nvc = nsteps > 0 ? (nsteps + m - 1) / m : 1;
  w.nvc = nvc;
  w.m_steps = m;
  size_t off = w.base.total;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  const size_t nv = (size_t)B * nvc;
  const size_t mat = (size_t)TC_D * TC_D * sizeof(float2);
  w.ops = take(nv * TC_N * TC_N * sizeof(float));
  w.ystart = take(nv * TC_D * sizeof(float2));
  w.lossv = take(nv * sizeof(double));
  w.rnv = take(nv * sizeof(float));
  if (save) {
    w.traj = take(nv * (size_t)(m + 1) * TC_D * sizeof(float2));
    w.sptraj = take(nv * (size_t)(m + 1) * TC_D * sizeof(float2));
    w.ev = take(nv * (size_t)(m + 1) * sizeof(float2));
    w.scales = take(nv * (size_t)(m / CH) * sizeof(float));
    w.G = take(nv * 3 * mat);
    w.gf = take(nv * TC_D * sizeof(float));
    w.lam0 = take(nv * TC_D * sizeof(float2));
    w.gAdir = take(nv * sizeof(double));
    w.lamend = take(nv * TC_D * sizeof(float2));
    w.wv = take(nv * sizeof(float));
    w.Gtot = take(3 * mat);
    w.gftot = take((size_t)TC_D * sizeof(float));
    w.lam0tot = take((size_t)TC_D * sizeof(float2));
  }
  w.total = off;
  return w;
}
}  // namespace

extern "C" {

size_t amps_psi_scan_workspace_bytes(int D, int B, int T, int save_for_bwd) {
  if (D <= 0 || D > 64 || B <= 0 || T < 1) return 0;
  return scan_ws_layout(B, T, save_for_bwd != 0).total;
}

int amps_psi_loss_fwd_scan(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                           float* loss_dev, void* ws_dev, size_t ws_bytes, int save_for_bwd,
                           void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (!p->psi0_dev) return fail(ctx, AMPS_E_INVALID, "psi0_dev is NULL");
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  if (B == 0) return AMPS_OK;
  if (!x_dev || !loss_dev || !ws_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  if (p->D > 64) return fail(ctx, AMPS_E_UNSUPPORTED, "the tensor-core scan supports D <= 64 (got %d)", p->D);
  const bool save = save_for_bwd != 0;
  const ScanWs L = scan_ws_layout(B, T, save);
  if (ws_bytes < L.total)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)ws_dev;
  const int DP = 64;   // every bond dimension is zero-padded to the UMMA shape
  rc = psi_prepare(ctx, p, DP, ws, L.base, T - 1, st);
  if (rc) return rc;
  const int nv = B * L.nvc;
  {
    const size_t smem = sizeof(ScanTcSmem) + 1024;
    auto kcomp = p->D <= 32 ? psi_compose_tc_kernel<true> : psi_compose_tc_kernel<false>;
    PROF_BEGIN(ctx, 2, st);
    kcomp<<<nv, TC_THREADS, smem, st>>>((const float2*)(ws + L.base.matN), (const float2*)(ws + L.base.matR),
                                        (const float2*)(ws + L.base.qtab), x_dev, T, aval(p), L.nvc,
                                        L.m_steps, (float*)(ws + L.ops));
    PROF_END(ctx, 2, st);
    LAUNCH_CHECK(ctx, "psi_compose_tc_kernel");
  }
  {
    const size_t bsm = (size_t)(2 * TC_N * TC_N + 3 * TC_N + 8) * sizeof(float);
    psi_scan_boundary_kernel<<<B, 256, bsm, st>>>((const float*)(ws + L.ops), (const float2*)(ws + L.base.psi0p),
                                                  L.nvc, (float2*)(ws + L.ystart), (float*)(ws + L.rnv));
  }
  LAUNCH_CHECK(ctx, "psi_scan_boundary_kernel");
  {
    auto kern = psi_fwd_uni_kernel<64, 4, true>;      // four lanes per row, as the batch kernels (launch_psi_fwd)
    const size_t smem = sizeof(FwdSmemUni<64, 4>);
    PROF_BEGIN(ctx, 0, st);
    kern<<<nv, 64 * 4, smem, st>>>((const float2*)(ws + L.base.matN), (const float2*)(ws + L.base.matR),
                                   (const float2*)(ws + L.base.matS), (const float2*)(ws + L.base.qtab),
                                   (const float2*)(ws + L.base.psi0p), x_dev, T, aval(p), (float*)nullptr,
                                   (double*)(ws + L.lossv), save ? (float2*)(ws + L.traj) : (float2*)nullptr,
                                   save ? (float*)(ws + L.scales) : (float*)nullptr, 0,
                                   (const float2*)(ws + L.ystart), L.nvc, L.m_steps,
                                   save ? (float2*)(ws + L.sptraj) : (float2*)nullptr,
                                   save ? (float2*)(ws + L.ev) : (float2*)nullptr, seg_full_f(T));
    PROF_END(ctx, 0, st);
    LAUNCH_CHECK(ctx, "psi_fwd_uni_kernel<virtual clips>");
  }
  psi_scan_sum_kernel<<<(B + 127) / 128, 128, 0, st>>>((const double*)(ws + L.lossv), B, L.nvc, loss_dev,
                                                       (double*)(ws + L.base.lossd));
  LAUNCH_CHECK(ctx, "psi_scan_sum_kernel");
  return AMPS_OK;
}

int amps_psi_loss_bwd_scan(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                           const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev,
                           void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  if (!grad_dev) return fail(ctx, AMPS_E_INVALID, "grad_dev is NULL");
  if (p->D > 64) return fail(ctx, AMPS_E_UNSUPPORTED, "the tensor-core scan supports D <= 64 (got %d)", p->D);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ng = amps_psi_grad_count(p->D);
  if (B == 0) {
    CUDA_TRY(ctx, cudaMemsetAsync(grad_dev, 0, ng * sizeof(float), st));
    return AMPS_OK;
  }
  if (!x_dev || !w_dev || !ws_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  const ScanWs L = scan_ws_layout(B, T, true);
  if (ws_bytes < L.total)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  char* ws = (char*)ws_dev;
  const int DP = 64;
  const int nv = B * L.nvc;
  const double cprime = -p->delta_t * (double)p->sigma * (double)p->sigma / 2.0;
  psi_scan_expand_w_kernel<<<(nv + 127) / 128, 128, 0, st>>>(w_dev, B, L.nvc, (float*)(ws + L.wv));
  LAUNCH_CHECK(ctx, "psi_scan_expand_w_kernel");
  auto kern_chain = psi_bwd_uni_kernel<64, 4, true, false>;   // pass 1: adjoint chain only
  auto kern_full = psi_bwd_uni_kernel<64, 4, true, true>;     // pass 2: chain + gradient tiles
  const size_t smem = sizeof(BwdSmemUni<64>);
  auto adjoint_pass = [&](const float2* lam_end) {
    auto kern = lam_end ? kern_full : kern_chain;
    kern<<<nv, 64 * 4, smem, st>>>((const float2*)(ws + L.base.matN), (const float2*)(ws + L.base.matRH),
                                   (const float2*)(ws + L.base.matS), (const float2*)(ws + L.base.qtab),
                                   (const float*)(ws + L.base.ttab), x_dev, T, aval(p), w_dev,
                                   (const float2*)(ws + L.traj), (const float*)(ws + L.scales), 0,
                                   (float2*)(ws + L.G), (float*)(ws + L.gf), (float2*)(ws + L.lam0),
                                   (double*)(ws + L.gAdir), lam_end, L.nvc, L.m_steps,
                                   (const float2*)(ws + L.sptraj), (const float2*)(ws + L.ev), seg_full_b(T),
                                   (float2*)nullptr);   // (function pointers carry no default argument)
  };
  PROF_BEGIN(ctx, 1, st);
  adjoint_pass(nullptr);                                   // d_j: chunk adjoints with a zero end condition
  LAUNCH_CHECK(ctx, "psi_bwd_uni_kernel<virtual clips, pass 1>");
  {
    const size_t bsm = (size_t)(2 * TC_N * TC_N + TC_N) * sizeof(float);
    psi_scan_boundary_bwd_kernel<<<B, 256, bsm, st>>>((const float*)(ws + L.ops), (const float*)(ws + L.rnv),
                                                      (const float2*)(ws + L.lam0), L.nvc,
                                                      (float2*)(ws + L.lamend));
    LAUNCH_CHECK(ctx, "psi_scan_boundary_bwd_kernel");
  }
  adjoint_pass((const float2*)(ws + L.lamend));            // true end conditions: tiles, g_f, g_A, Lam_0
  PROF_END(ctx, 1, st);
  LAUNCH_CHECK(ctx, "psi_bwd_uni_kernel<virtual clips, pass 2>");
  {
    const int total = 3 * DP * DP + 2 * DP;
    psi_reduce_clips_kernel<<<(total + 127) / 128, 128, 0, st>>>(
        (const float2*)(ws + L.G), (const float*)(ws + L.gf), (const float2*)(ws + L.lam0), nv, DP,
        (float2*)(ws + L.Gtot), (float*)(ws + L.gftot), (float2*)(ws + L.lam0tot), L.nvc);
    LAUNCH_CHECK(ctx, "psi_reduce_clips_kernel");
    psi_grad_finalize_kernel<<<p->D, 128, 0, st>>>(
        (const float2*)(ws + L.Gtot), (const float*)(ws + L.gftot), (const float2*)(ws + L.lam0tot),
        (const double*)(ws + L.gAdir), (const double*)(ws + L.lossv), (const float*)(ws + L.wv), nv,
        (const float2*)(ws + L.base.matR), p->D, DP, cprime, aval(p), grad_dev);
    LAUNCH_CHECK(ctx, "psi_grad_finalize_kernel");
  }
  return AMPS_OK;
}

size_t amps_psi_sample_workspace_bytes(int D, int L, int n) {
  const int DP = padded_dim(D);
  if (DP < 0 || L < 0 || n < 0) return 0;
  // step operators, t_k / q_k tables + the noise tensor transposed to [n][L]
  return psi_ws_layout(DP, 0, L, 0, false).total + align_up((size_t)n * L * sizeof(float));
}

int amps_psi_sample(amps_ctx* ctx, const amps_params* p, const float* noise_dev, int L_, int n,
                    float* out_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (!p->psi0_dev) return fail(ctx, AMPS_E_INVALID, "psi0_dev is NULL");
  if (L_ < 0 || n < 0) return fail(ctx, AMPS_E_INVALID, "bad shape L=%d n=%d", L_, n);
  if (L_ == 0 || n == 0) return AMPS_OK;
  if (!noise_dev || !out_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  const int DP = padded_dim(p->D);
  if (DP < 0) return fail(ctx, AMPS_E_UNSUPPORTED, "bond dimension %d > 128 not supported", p->D);
  const PsiWs L = psi_ws_layout(DP, 0, L_, 0, false);
  const size_t need = L.total + align_up((size_t)n * L_ * sizeof(float));
  if (!ws_dev || ws_bytes < need)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)ws_dev;
  rc = psi_prepare(ctx, p, DP, ws, L, L_, st);
  if (rc) return rc;
  float* noiseT = (float*)(ws + L.total);
  prep_transpose_kernel<<<dim3((n + 31) / 32, (L_ + 31) / 32), dim3(32, 8), 0, st>>>(noise_dev, L_, n, noiseT);
  LAUNCH_CHECK(ctx, "prep_transpose_kernel");
  noise_dev = noiseT;
  if (DP == 128) {   // rows of N and R split over a 4-CTA cluster per waveform
    PROF_BEGIN(ctx, 2, st);
    CUDA_TRY(ctx, launch_cluster(psi_sample_c4_kernel<128, C4_CL>, n, C4_CL, 512, sizeof(SampleC4Smem<128, C4_CL>), st,
                                 (const float2*)(ws + L.matN), (const float2*)(ws + L.matR), (const float2*)(ws + L.qtab),
                                 (const float2*)(ws + L.psi0p), noise_dev, L_, n, p->A, (float)p->delta_t, out_dev));
    PROF_END(ctx, 2, st);
    LAUNCH_CHECK(ctx, "psi_sample_c4_kernel");
    return AMPS_OK;
  }
  return dispatch_dp(DP, [&](auto dp, auto nq) -> int {
    constexpr int DPc = decltype(dp)::value;
    constexpr int NQc = DPc == 64 ? 4 : decltype(nq)::value;   // D = 64: four lanes per row, as the training kernels (4.4)
    auto kern = psi_sample_kernel<DPc, NQc>;
    const size_t smem = sizeof(SampleSmem<DPc>);
    PROF_BEGIN(ctx, 2, st);
    kern<<<n, DPc * NQc, smem, st>>>((const float2*)(ws + L.matN), (const float2*)(ws + L.matR),
                                     (const float2*)(ws + L.qtab), (const float2*)(ws + L.psi0p),
                                     noise_dev, L_, n, p->A, (float)p->delta_t, out_dev);
    PROF_END(ctx, 2, st);
    LAUNCH_CHECK(ctx, "psi_sample_kernel");
    return AMPS_OK;
  });
}

int amps_psi_evolve(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                    float* traj_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!ctx) return AMPS_E_INVALID;
  if (!traj_dev && B > 0 && T > 1) return fail(ctx, AMPS_E_INVALID, "traj_dev is NULL");
  if (B <= 0 || T <= 1) return (B < 0 || T < 1) ? fail(ctx, AMPS_E_INVALID, "bad shape") : AMPS_OK;
  const int DP = padded_dim(p ? p->D : 0);
  if (DP < 0) return fail(ctx, AMPS_E_UNSUPPORTED, "bond dimension not supported");
  const PsiWs L = psi_ws_layout(DP, B, T - 1, T, true);
  if (ws_bytes < L.total)
    return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  // per-clip loss lands in the (unused) gf slot of the workspace
  char* ws = (char*)ws_dev;
  int rc = amps_psi_loss_fwd(ctx, p, x_dev, B, T, (float*)(ws + L.gf), ws_dev, ws_bytes, 1, stream);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t items = (size_t)B * (T - 1);
  int blocks = (int)((items * 32 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  psi_lab_traj_kernel<<<blocks, 256, 0, st>>>((const float2*)(ws + L.traj), p->freqs_dev, (const float*)(ws + L.ttab),
                                              B, T, p->D, DP, (float2*)traj_dev);
  LAUNCH_CHECK(ctx, "psi_lab_traj_kernel");
  return AMPS_OK;
}

int amps_psi_loss_grad_host(amps_ctx* ctx, const amps_host_params* hp, const float* x_host, int B,
                            int T, float w, float* loss_host, float* grad_host) {
  if (!ctx) return AMPS_E_INVALID;
  if (!hp || !x_host || !loss_host || !grad_host) return fail(ctx, AMPS_E_INVALID, "NULL argument");
  if (B <= 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  const int D = hp->D;
  if (padded_dim(D) < 0) return fail(ctx, AMPS_E_UNSUPPORTED, "bond dimension %d not supported", D);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!ctx->hstream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->hstream, cudaStreamNonBlocking));
  cudaStream_t st = ctx->hstream;
  const size_t ng = amps_psi_grad_count(D);
  const size_t wsb = amps_psi_workspace_bytes(D, B, T, 1);
  // layout of the host-entry device buffer
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  const size_t oR = take((size_t)D * D * 8), oF = take((size_t)D * 4), oP = take((size_t)D * 8);
  const size_t oX = take((size_t)B * T * 4), oL = take((size_t)B * 4), oW = take((size_t)B * 4);
  const size_t oG = take(ng * 4), oWs = take(wsb);
  if (ctx->hbuf_bytes < off) {
    if (ctx->hbuf) CUDA_TRY(ctx, cudaFree(ctx->hbuf));
    ctx->hbuf = nullptr;
    ctx->hbuf_bytes = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->hbuf, off));
    ctx->hbuf_bytes = off;
  }
  char* d = (char*)ctx->hbuf;
  CUDA_TRY(ctx, cudaMemcpyAsync(d + oR, hp->R, (size_t)D * D * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(d + oF, hp->freqs, (size_t)D * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(d + oP, hp->psi0, (size_t)D * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(d + oX, x_host, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
  fill_kernel<<<(B + 255) / 256, 256, 0, st>>>((float*)(d + oW), B, w);
  LAUNCH_CHECK(ctx, "fill_kernel");
  amps_params p{};
  p.D = D;
  p.R_dev = (const float*)(d + oR);
  p.freqs_dev = (const float*)(d + oF);
  p.psi0_dev = (const float*)(d + oP);
  p.A = hp->A;
  p.sigma = hp->sigma;
  p.delta_t = hp->delta_t;
  int rc = amps_psi_loss_fwd(ctx, &p, (const float*)(d + oX), B, T, (float*)(d + oL), d + oWs, wsb, 1, st);
  if (rc) return rc;
  rc = amps_psi_loss_bwd(ctx, &p, (const float*)(d + oX), B, T, (const float*)(d + oW), d + oWs, wsb,
                         (float*)(d + oG), st);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(loss_host, d + oL, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(grad_host, d + oG, ng * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  return AMPS_OK;
}

// ---- RhoCMPS -----------------------------------------------------------------------------
size_t amps_rho_workspace_bytes(int D, int B, int T, int save_for_bwd) {
  return rho_workspace_bytes(D, B, T, save_for_bwd != 0);
}
size_t amps_rho_grad_count(int D) { return D > 0 ? (size_t)4 * D * D + (size_t)D + 2 : 0; }

int amps_rho_loss_fwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      float* loss_dev, void* ws_dev, size_t ws_bytes, int save_for_bwd, void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (!p->rho0_dev) return fail(ctx, AMPS_E_INVALID, "rho0_dev is NULL");
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  if (B == 0) return AMPS_OK;
  if (!x_dev || !loss_dev || !ws_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  if (p->D > RHO_MAX_D) return fail(ctx, AMPS_E_UNSUPPORTED, "rho kernels support D <= %d", RHO_MAX_D);
  const bool save = save_for_bwd != 0;
  const RhoWs L = rho_ws_layout(p->D, B, T, save);
  if (ws_bytes < L.total) return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)ws_dev;
  rc = rho_tables(ctx, p, T - 1, false, ws, L, st);
  if (rc) return rc;
  rc = rho_launch_data(p, (const float*)(ws + L.ttab), (const float2*)(ws + L.qtab), nullptr, x_dev, B, T, loss_dev, nullptr,
                       save ? (float2*)(ws + L.ftraj) : nullptr, (double*)(ws + L.lossd), st);
  if (rc) return fail(ctx, rc, "rho kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  ctx->launches++;
  return AMPS_OK;
}

int amps_rho_loss_bwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev, void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  if (!grad_dev) return fail(ctx, AMPS_E_INVALID, "grad_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    CUDA_TRY(ctx, cudaMemsetAsync(grad_dev, 0, amps_rho_grad_count(p->D) * sizeof(float), st));
    return AMPS_OK;
  }
  if (!x_dev || !w_dev || !ws_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  if (p->D > RHO_MAX_D) return fail(ctx, AMPS_E_UNSUPPORTED, "rho kernels support D <= %d", RHO_MAX_D);
  const RhoWs L = rho_ws_layout(p->D, B, T, true);
  if (ws_bytes < L.total) return fail(ctx, AMPS_E_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, L.total);
  // t_k and q_k are the ones the saving forward left in this workspace
  rc = rho_launch_bwd(p, (const float*)((char*)ws_dev + L.ttab), (const float2*)((char*)ws_dev + L.qtab), x_dev, B, T,
                      w_dev, (char*)ws_dev, L, grad_dev, st);
  if (rc) return fail(ctx, rc, "rho backward launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  ctx->launches += 2;
  return AMPS_OK;
}

int amps_rho_evolve(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                    float* traj_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (!p->rho0_dev) return fail(ctx, AMPS_E_INVALID, "rho0_dev is NULL");
  if (B < 0 || T < 1) return fail(ctx, AMPS_E_INVALID, "bad shape B=%d T=%d", B, T);
  if (B == 0 || T == 1) return AMPS_OK;
  if (!x_dev || !traj_dev) return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  if (p->D > RHO_MAX_D) return fail(ctx, AMPS_E_UNSUPPORTED, "rho kernels support D <= %d", RHO_MAX_D);
  const RhoWs L = rho_ws_layout(p->D, B, T, false);
  if (!ws_dev || ws_bytes < L.total) return fail(ctx, AMPS_E_WORKSPACE, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)ws_dev;
  rc = rho_tables(ctx, p, T - 1, true, ws, L, st);
  if (rc) return rc;
  rc = rho_launch_data(p, (const float*)(ws + L.ttab), (const float2*)(ws + L.qtab), (const float2*)(ws + L.ptab), x_dev, B,
                       T, nullptr, (float2*)traj_dev, nullptr, nullptr, st);
  if (rc) return fail(ctx, rc, "rho kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  ctx->launches++;
  return AMPS_OK;
}

int amps_rho_sample(amps_ctx* ctx, const amps_params* p, const float* noise_dev, int L, int n,
                    float* out_dev, float* traj_dev, float* purity_dev, void* ws_dev,
                    size_t ws_bytes, void* stream) {
  int rc = check_common(ctx, p);
  if (rc) return rc;
  if (!p->rho0_dev) return fail(ctx, AMPS_E_INVALID, "rho0_dev is NULL");
  if (L < 0 || n < 0) return fail(ctx, AMPS_E_INVALID, "bad shape L=%d n=%d", L, n);
  if (L == 0 || n == 0) return AMPS_OK;
  if (!noise_dev) return fail(ctx, AMPS_E_INVALID, "noise_dev is NULL");
  if (p->D > RHO_MAX_D) return fail(ctx, AMPS_E_UNSUPPORTED, "rho kernels support D <= %d", RHO_MAX_D);
  const RhoWs Lw = rho_ws_layout(p->D, n, L + 1, false);
  if (!ws_dev || ws_bytes < Lw.total) return fail(ctx, AMPS_E_WORKSPACE, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)ws_dev;
  rc = rho_tables(ctx, p, L, traj_dev != nullptr, ws, Lw, st);
  if (rc) return rc;
  rc = rho_launch_sample(p, (const float*)(ws + Lw.ttab), (const float2*)(ws + Lw.qtab),
                         traj_dev ? (const float2*)(ws + Lw.ptab) : nullptr, noise_dev, L, n, out_dev,
                         (float2*)traj_dev, purity_dev, ws_dev, st);
  if (rc) return fail(ctx, rc, "rho kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  ctx->launches++;
  return AMPS_OK;
}

// ---- a1 / a2: raw variables <-> effective parameters, regulariser --------------------------------
int amps_psi_params_fwd(amps_ctx* ctx, int D, const float* Rx_dev, const float* Ry_dev,
                        const float* freqs_raw_dev, const float* psi_x_dev, const float* psi_y_dev,
                        float r_scale, float f_scale, float h_reg, float r_reg, float* R_eff_dev,
                        float* freqs_eff_dev, float* psi0_dev, float* aux_dev, void* stream) {
  if (!ctx) return AMPS_E_INVALID;
  if (D <= 0) return fail(ctx, AMPS_E_INVALID, "bond dimension D=%d must be positive", D);
  if (!Rx_dev || !Ry_dev || !freqs_raw_dev || !psi_x_dev || !psi_y_dev || !R_eff_dev || !freqs_eff_dev ||
      !psi0_dev || !aux_dev)
    return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  psi_params_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(Rx_dev, Ry_dev, freqs_raw_dev, psi_x_dev, psi_y_dev, D,
                                                             r_scale, f_scale, h_reg, r_reg, (float2*)R_eff_dev,
                                                             freqs_eff_dev, (float2*)psi0_dev, aux_dev);
  LAUNCH_CHECK(ctx, "psi_params_fwd_kernel");
  return AMPS_OK;
}

int amps_psi_params_bwd(amps_ctx* ctx, int D, const float* Rx_dev, const float* Ry_dev,
                        const float* freqs_raw_dev, const float* psi_x_dev, const float* psi_y_dev,
                        float r_scale, float f_scale, float h_reg, float r_reg, const float* aux_dev,
                        const float* gR_dev, const float* gfreqs_dev, const float* gpsi0_dev,
                        const float* greg_dev, float* gRx_dev, float* gRy_dev, float* gfreqs_raw_dev,
                        float* gpsi_x_dev, float* gpsi_y_dev, void* stream) {
  if (!ctx) return AMPS_E_INVALID;
  if (D <= 0) return fail(ctx, AMPS_E_INVALID, "bond dimension D=%d must be positive", D);
  if (!Rx_dev || !Ry_dev || !freqs_raw_dev || !psi_x_dev || !psi_y_dev || !aux_dev || !gR_dev || !gfreqs_dev ||
      !gpsi0_dev || !gRx_dev || !gRy_dev || !gfreqs_raw_dev || !gpsi_x_dev || !gpsi_y_dev)
    return fail(ctx, AMPS_E_INVALID, "NULL buffer");
  psi_params_bwd_kernel<<<1, 256, (size_t)D * sizeof(float2), (cudaStream_t)stream>>>(
      Rx_dev, Ry_dev, freqs_raw_dev, psi_x_dev, psi_y_dev, D, r_scale, f_scale, h_reg, r_reg, aux_dev,
      (const float2*)gR_dev, gfreqs_dev, (const float2*)gpsi0_dev, greg_dev, gRx_dev, gRy_dev, gfreqs_raw_dev,
      gpsi_x_dev, gpsi_y_dev);
  LAUNCH_CHECK(ctx, "psi_params_bwd_kernel");
  return AMPS_OK;
}

// ---- data-parallel communicator over NCCL (NVLink / NVSwitch) --------------------------------
int amps_comm_unique_id(void* id_out) {
  if (!id_out) return AMPS_E_INVALID;
  NcclApi* n = nccl_api();
  if (!n) return AMPS_E_UNSUPPORTED;
  NcclId id;
  if (n->GetUniqueId(&id) != 0) return AMPS_E_CUDA;
  memcpy(id_out, &id, sizeof(id));
  return AMPS_OK;
}

int amps_comm_init(amps_ctx* ctx, const void* id, int rank, int nranks) {
  if (!ctx) return AMPS_E_INVALID;
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, AMPS_E_INVALID, "bad communicator arguments");
  if (ctx->nccl_comm) return fail(ctx, AMPS_E_STATE, "communicator already initialised");
  NcclApi* n = nccl_api();
  if (!n) return fail(ctx, AMPS_E_UNSUPPORTED, "libnccl.so.2 not found (set AMPS_NCCL_LIB)");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  void* comm = nullptr;
  const int rc = n->CommInitRank(&comm, nranks, nid, rank);
  if (rc != 0) return fail(ctx, AMPS_E_CUDA, "ncclCommInitRank: %s", n->GetErrorString ? n->GetErrorString(rc) : "error");
  ctx->nccl_comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_size = nranks;
  return AMPS_OK;
}

int amps_allreduce_grads(amps_ctx* ctx, float* packed_dev, size_t count, void* stream) {
  if (!ctx) return AMPS_E_INVALID;
  if (!ctx->nccl_comm) return fail(ctx, AMPS_E_STATE, "amps_comm_init has not been called");
  if (!packed_dev && count) return fail(ctx, AMPS_E_INVALID, "packed_dev is NULL");
  if (count == 0) return AMPS_OK;
  NcclApi* n = nccl_api();
  const int rc = n->AllReduce(packed_dev, packed_dev, count, kNcclFloat, kNcclSum, ctx->nccl_comm, (cudaStream_t)stream);
  if (rc != 0) return fail(ctx, AMPS_E_CUDA, "ncclAllReduce: %s", n->GetErrorString ? n->GetErrorString(rc) : "error");
  return AMPS_OK;
}

int amps_comm_destroy(amps_ctx* ctx) {
  if (!ctx) return AMPS_E_INVALID;
  if (ctx->nccl_comm) {
    NcclApi* n = nccl_api();
    if (n) n->CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->comm_size = 1;
    ctx->comm_rank = 0;
  }
  return AMPS_OK;
}

}  // extern "C"
