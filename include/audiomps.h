/*
 * audiomps.h -- C ABI of libaudiomps.so, the B200-native (sm_100a) implementation of the
 * continuous-MPS ("AudioMPS") time-step scan of AustenLamacraft/audio-mps.
 *
 * The reference has no FFI: the boundary it exposes for this path is the Python model API of
 * /root/reference/model.py (PsiCMPS / RhoCMPS).  Each entry point below names the reference
 * interface it replaces (file:line into the reference tree).  The Python host package
 * (audio_mps_b200/) binds these with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative AMPS_E_* code otherwise; no C++ exception
 *     crosses the ABI; amps_last_error() returns a per-context message.
 *   - all `*_dev` pointers are CUDA device pointers on the context's device; the caller owns them
 *     (including the workspace).  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  Device entry points never synchronise; host entry points (`*_host`) synchronise
 *     the context's private stream before returning.
 *   - complex64 arrays are interleaved (re, im) float pairs, row-major.
 *   - parameters handed to the library are the EFFECTIVE ones (after model.py:31-52 and
 *     model.py:221-222 / 127-130): R_eff[i,j] = R[i,j] - R[j,j], freqs, normalised psi_0 / rho_0.
 *     The raw->effective chain (O(D^2), off the hot path) stays in host code.
 *   - time is the float32 running sum t_{k+1} = fl32(t_k + fl32(delta_t)) (model.py:16,157,281);
 *     the library reproduces it bit-exactly and forms phases from fl32(f_c * t_k).
 */
#ifndef AUDIOMPS_H_
#define AUDIOMPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMPS_VERSION 200 /* 0.2.0 */

enum {
  AMPS_OK = 0,
  AMPS_E_INVALID = -1,     /* bad argument (null pointer, negative size, ...)          */
  AMPS_E_UNSUPPORTED = -2, /* bond dimension / shape outside what the kernels cover    */
  AMPS_E_WORKSPACE = -3,   /* workspace too small (see amps_*_workspace_bytes)          */
  AMPS_E_CUDA = -4,        /* a CUDA runtime call failed; see amps_last_error           */
  AMPS_E_STATE = -5        /* call order violated (e.g. bwd without a saving fwd)       */
};

typedef struct amps_ctx amps_ctx;

/* Effective parameters of one model (device pointers + host scalars).
 * Replaces the tensors CMPS.__init__ / PsiCMPS.__init__ / RhoCMPS.__init__ build:
 * model.py:9-52 (R, freqs, A, sigma, delta_t), :221-222 (psi_0), :127-130 (rho_0). */
typedef struct amps_params {
  int32_t D;            /* bond dimension (hparams.bond_dim, model.py:11)                   */
  int32_t reserved;
  const float* R_dev;     /* complex64 [D,D]  effective R (model.py:41-42)                  */
  const float* freqs_dev; /* float32   [D]    effective freqs (model.py:44-50)              */
  const float* psi0_dev;  /* complex64 [D]    normalised psi_0 (model.py:221-222); Psi only */
  const float* rho0_dev;  /* complex64 [D,D]  rho_0 (model.py:127-130); Rho only            */
  float A;                /* model.py:19 (trainable scalar, value at this step)             */
  float sigma;            /* model.py:21                                                    */
  double delta_t;         /* model.py:15 (python double; dt32 = (float)delta_t, model.py:16)*/
  const float* A_dev;     /* optional: device float32 holding A; when non-NULL the Psi loss /
                           * gradient / scan entry points read A from it and ignore `A`, so a
                           * training loop never reads the parameter back to the host (the
                           * sampler and the Rho entry points use the host value)          */
} amps_params;

/* ---- lifecycle ------------------------------------------------------------------------- */
int amps_version(void);
int amps_create(int device, amps_ctx** out);
int amps_destroy(amps_ctx* ctx);
const char* amps_last_error(const amps_ctx* ctx);
/* number of kernels the context has launched so far (bench.py's gpu_launches evidence) */
int64_t amps_launch_count(const amps_ctx* ctx);

/* The float32 running-sum time table t_0 = 0, t_{k+1} = fl32(t_k + fl32(delta_t)) of model.py:16,157,281,
 * out[0..n), computed on the host by the same piecewise-exact generator the device entry points use
 * (bit-identical to the sequential sum; tests/test_host_logic.py). */
int amps_time_table_host(double delta_t, int n, float* out);

/* ---- measurement helpers (bench.py) ------------------------------------------------------ */
/* When enabled, the context brackets the dominant kernel of each entry point with CUDA events on
 * the launch stream; amps_get_kernel_ms(which: 0 = psi forward, 1 = psi backward (all of its kernels),
 * 2 = psi sampler / operator composition / the tensor-core expectation pass inside 0,
 * 3 = the tensor-core gradient-tile kernels inside 1)
 * synchronises on the stop event and returns the last launch's duration. */
int amps_set_profiling(amps_ctx* ctx, int enable);
int amps_get_kernel_ms(amps_ctx* ctx, int which, float* ms);
/* FP32 FMA issue-rate microbenchmark on the context's device (dense FFMA chains, best of 5). */
double amps_fma_peak_tflops(amps_ctx* ctx);
/* packed != 0: the same chains issued as FFMA2 (fma.rn.f32x2). */
double amps_fma_peak_tflops2(amps_ctx* ctx, int packed);

/* ---- PsiCMPS --------------------------------------------------------------------------- */

/* Bond dimensions: loss, gradient and trajectory 1 <= D <= 128 (zero-padded to 8/16/32/64/128; above
 * 64 the matrices are row-split over a 4-CTA cluster), sampler likewise; tensor-core scan D <= 64; Rho
 * D <= 32.  Anything else returns AMPS_E_UNSUPPORTED (and a workspace size of 0).
 *
 * bytes of caller-owned workspace for the Psi loss forward/backward at (D, B clips, T samples).
 * save_for_bwd != 0 adds the state trajectory the adjoint sweep consumes. */
size_t amps_psi_workspace_bytes(int D, int B, int T, int save_for_bwd);

/* Per-clip negative log-likelihood loss_b, the fold of PsiCMPS._build_loss_psi /
 * _psi_and_loss_update (model.py:257-267, 276-282, 293-334) BEFORE the reduce_mean (:267).
 *   x_dev    float32 [B,T] row-major waveform (data_iterator)
 *   loss_dev float32 [B]
 * With save_for_bwd the workspace keeps what amps_psi_loss_bwd needs. */
int amps_psi_loss_fwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      float* loss_dev, void* ws_dev, size_t ws_bytes, int save_for_bwd,
                      void* stream);

/* Gradient of L = sum_b w_b * loss_b with respect to the effective parameters; replaces the
 * reverse while_loop tf.gradients builds for train.py:89.  Must follow a saving forward on the
 * same (params, x, workspace).
 *   w_dev    float32 [B]   per-clip weights (reference: 1/B, model.py:267)
 *   grad_dev float32 [2*D*D + 3*D + 2] packed:
 *            gR (complex64 [D,D], dL/dRe + i dL/dIm) | gfreqs [D] | gpsi0 (complex64 [D]) |
 *            gA [1] | sum_b w_b*loss_b [1]
 * (the packed buffer is what one NCCL all-reduce sums across data-parallel ranks). */
int amps_psi_loss_bwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev,
                      void* stream);
size_t amps_psi_grad_count(int D); /* = 2*D*D + 3*D + 2 */

/* Checkpointed loss / gradient: the state is kept every K steps only (BASELINE north_star: "adjoint backward
 * that recomputes from state checkpoints every K steps"; the reference's own memory wall is the O(T)
 * activation stack of tf.foldl's reverse loop, model.py:265-266, train.py:91 "Unrolling in time?").
 *   K == 1 : identical to amps_psi_loss_fwd(save_for_bwd = 1) / amps_psi_loss_bwd -- the whole trajectory
 *            (12 + 16*DP bytes per step and clip) stays in the workspace; fastest when it fits.
 *   K  > 1 : the forward stores ONE state (8*DP bytes) per K steps; the backward walks the windows from the
 *            last to the first, re-running the forward kernel over window j-1 (on a second stream owned by
 *            the context) while the adjoint kernel sweeps window j.  Workspace: checkpoints + two
 *            window-sized trajectory buffers.  K is rounded up to a whole number of rescale chunks
 *            (amps_psi_ckpt_interval: 32 steps, 16 above D = 64).  Every window costs two kernel launches,
 *            so K of a few thousand steps is the useful range (K = 2048: 34 MB instead of 2.1 GB per 64
 *            clips at D = 32).
 * Same results as K == 1 (the replay is the same kernel from the same state: bit-identical trajectory);
 * loss_dev, w_dev, grad_dev as above.  The backward must follow the forward on the same (params, x,
 * workspace, K). */
int amps_psi_ckpt_interval(int D, int K);
size_t amps_psi_workspace_bytes_k(int D, int B, int T, int K);
int amps_psi_loss_fwd_k(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T, int K,
                        float* loss_dev, void* ws_dev, size_t ws_bytes, void* stream);
int amps_psi_loss_bwd_k(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T, int K,
                        const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev, void* stream);

/* Parallel-in-time loss (and gradient) for SMALL batches (same results as amps_psi_loss_fwd / _bwd;
 * D <= 64, zero padded to 64): the clip is cut into ~#SMs/B time chunks whose step operators are
 * composed on the tcgen05 tensor cores (complex DxD as real 2Dx2D, kind::tf32 with a 3-pass hi/lo
 * split, running product in tensor memory), a short sequential pass over the chunk operators gives
 * the chunk start states, and every chunk is then replayed in parallel by the sequential kernel as
 * a "virtual clip".  8 D^3 instead of 24 D^2 flops per step: use it only when B is far below the SM
 * count.  With save_for_bwd != 0 the replay keeps its trajectories and amps_psi_loss_bwd_scan runs
 * the adjoint the same way: per-chunk adjoints with a zero end condition, a sequential adjoint pass
 * over the chunk operators (Lam_j = d_j + C_j^dag Lam_{j+1} / |C_j y_j|), then per-chunk adjoints from
 * the true end conditions, all chunks in parallel.  grad_dev as for amps_psi_loss_bwd. */
size_t amps_psi_scan_workspace_bytes(int D, int B, int T, int save_for_bwd);
int amps_psi_loss_fwd_scan(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                           float* loss_dev, void* ws_dev, size_t ws_bytes, int save_for_bwd,
                           void* stream);
int amps_psi_loss_bwd_scan(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                           const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev,
                           void* stream);

/* PsiCMPS.sample / _psi_and_sample_update (model.py:242-251, 284-291) with the noise tensor
 * supplied by the caller (the reference draws it once, model.py:246).
 *   noise_dev float32 [L,n] (time-major, as tf.random_normal([length, num_samples]))
 *   out_dev   float32 [n,L] = A * cumulative X_t (model.py:251)
 *   ws_dev    caller-owned workspace of amps_psi_sample_workspace_bytes(D, L, n) bytes (step
 *             operators, t_k and q_k tables of this call) */
size_t amps_psi_sample_workspace_bytes(int D, int L, int n);
int amps_psi_sample(amps_ctx* ctx, const amps_params* p, const float* noise_dev, int L, int n,
                    float* out_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* PsiCMPS.psi_evolve_with_data / _psi_update (model.py:231-240, 269-274):
 *   traj_dev complex64 [B, T-1, D] normalised lab-frame psi after every step.
 * Uses the same workspace size as a saving forward. */
int amps_psi_evolve(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                    float* traj_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* Host-buffer training step (bench.py "e2e"): copies x (pinned or pageable host memory) to the
 * device, runs forward + backward with w_b = 1/B_global, copies loss[B] and the packed gradient
 * back.  Parameters are HOST arrays here.  Device scratch is owned by the context and grown on
 * demand.  Synchronises before returning. */
typedef struct amps_host_params {
  int32_t D;
  int32_t reserved;
  const float* R;     /* complex64 [D,D] host */
  const float* freqs; /* float32 [D] host     */
  const float* psi0;  /* complex64 [D] host   */
  float A;
  float sigma;
  double delta_t;
} amps_host_params;
int amps_psi_loss_grad_host(amps_ctx* ctx, const amps_host_params* p, const float* x_host, int B,
                            int T, float w, float* loss_host, float* grad_host);

/* ---- RhoCMPS --------------------------------------------------------------------------- */

size_t amps_rho_workspace_bytes(int D, int B, int T, int save_for_bwd);

/* RhoCMPS._build_loss_rho / _rho_and_loss_update (model.py:132-142, 152-158, 169-203):
 * per-clip loss before the reduce_mean.  With save_for_bwd the workspace keeps the interaction-frame
 * density matrix at the start of every step (B*T*D*D complex64) for amps_rho_loss_bwd. */
int amps_rho_loss_fwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      float* loss_dev, void* ws_dev, size_t ws_bytes, int save_for_bwd, void* stream);

/* Gradient of sum_b w_b loss_b wrt the effective parameters of RhoCMPS (replaces tf.gradients for
 * train.py:89 with --mps_model=rho_mps).  grad_dev float32 [4*D*D + D + 2] packed:
 *   gR (complex64 [D,D]) | gfreqs [D] | grho0 (complex64 [D,D]) | gA [1] | sum_b w_b*loss_b [1] */
int amps_rho_loss_bwd(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                      const float* w_dev, void* ws_dev, size_t ws_bytes, float* grad_dev, void* stream);
size_t amps_rho_grad_count(int D); /* = 4*D*D + D + 2 */

/* RhoCMPS.rho_evolve_with_data (model.py:76-85): traj_dev complex64 [B, T-1, D, D]. */
int amps_rho_evolve(amps_ctx* ctx, const amps_params* p, const float* x_dev, int B, int T,
                    float* traj_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* RhoCMPS.sample / rho_evolve_with_sampling / purity (model.py:87-112, 160-167) from a supplied
 * noise tensor [L,n].  Any of the three outputs may be NULL.
 *   out_dev     float32   [n,L]      A * cumulative X_t
 *   traj_dev    complex64 [n,L,D,D]  rho after every step
 *   purity_dev  float32   [n,L]      Re tr(rho_k^2) */
int amps_rho_sample(amps_ctx* ctx, const amps_params* p, const float* noise_dev, int L, int n,
                    float* out_dev, float* traj_dev, float* purity_dev, void* ws_dev,
                    size_t ws_bytes, void* stream);

/* CMPS.__init__ / PsiCMPS.__init__ parameterisation (model.py:36-50, 218-222, 327-334) and the
 * regulariser of train.py:55-60 in one launch, and its reverse chain -- so that a training step does
 * not spend ~60 framework launches on O(D^2) bookkeeping.
 *   R_eff[i,j] = r_scale*(Rx+iRy)[i,j] - r_scale*(Rx+iRy)[j,j]   (r_scale = rsqrt(r_reg), or 1 with R_in)
 *   freqs_eff  = f_scale*freqs_raw                                (f_scale = rsqrt(h_reg), or 1 with freqs_in)
 *   psi0       = (psi_x+i psi_y) * rsqrt(max(sum|.|^2, 1e-12))
 *   aux_dev float32[4]: [0] = h_reg*sum freqs_eff^2 + r_reg*sum |R_eff|^2, [1..2] = state for the backward
 * Backward: gR/gpsi0 complex64 (dL/dRe + i dL/dIm), gfreqs float32, greg_dev = dL/d aux[0] (device scalar,
 * may be NULL = 0); outputs are the gradients wrt the raw variables. */
int amps_psi_params_fwd(amps_ctx* ctx, int D, const float* Rx_dev, const float* Ry_dev,
                        const float* freqs_raw_dev, const float* psi_x_dev, const float* psi_y_dev,
                        float r_scale, float f_scale, float h_reg, float r_reg, float* R_eff_dev,
                        float* freqs_eff_dev, float* psi0_dev, float* aux_dev, void* stream);
int amps_psi_params_bwd(amps_ctx* ctx, int D, const float* Rx_dev, const float* Ry_dev,
                        const float* freqs_raw_dev, const float* psi_x_dev, const float* psi_y_dev,
                        float r_scale, float f_scale, float h_reg, float r_reg, const float* aux_dev,
                        const float* gR_dev, const float* gfreqs_dev, const float* gpsi0_dev,
                        const float* greg_dev, float* gRx_dev, float* gRy_dev, float* gfreqs_raw_dev,
                        float* gpsi_x_dev, float* gpsi_y_dev, void* stream);

/* ---- data parallelism: ONE all-reduce of the packed gradient per step ------------------------
 * Batch is the only shard axis (model.py:258-267): every rank runs amps_psi_loss_fwd/_bwd on its clips
 * with w_b = 1/B_global and sums the packed buffer [gR | gfreqs | gpsi0 | gA | sum_b w_b loss_b]
 * (2*D*D + 3*D + 2 floats) across ranks.  The communicator is NCCL (NVLink 5 / NVSwitch), resolved at
 * run time from the libnccl.so.2 of the process (no link-time dependency; AMPS_NCCL_LIB overrides).
 *   amps_comm_unique_id: rank 0 fills AMPS_COMM_ID_BYTES bytes (an ncclUniqueId) and hands them to the
 *                        other ranks out of band (file, socket, torch.distributed, MPI ...)
 *   amps_comm_init:      collective over the nranks contexts (one per GPU / process)
 *   amps_allreduce_grads: in-place float32 sum on `stream`
 * A host that already has a communicator (torch.distributed) may all-reduce the buffer itself. */
#define AMPS_COMM_ID_BYTES 128
int amps_comm_unique_id(void* id_out);
int amps_comm_init(amps_ctx* ctx, const void* id, int rank, int nranks);
int amps_allreduce_grads(amps_ctx* ctx, float* packed_dev, size_t count, void* stream);
int amps_comm_destroy(amps_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOMPS_H_ */
