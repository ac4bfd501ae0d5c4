"""Round-2 goldens (test infrastructure), minted with the float64 C/OpenMP restatement
(oracle/cmps_ref.c -- lab frame, per-step normalisation; independent of the CUDA chain form):

  psi_c2_sample_full      BASELINE config[2] at FULL size: D=32, 256 samples x 64000 steps from a fixed
                          noise tensor (seed recorded).  The 256 x 64000 float64 output is 131 MB, so
                          every 125th column is stored (256 x 512) -- the output is a CUMULATIVE sum,
                          any drift of the fed-back state shows in every later column -- plus row sums.
  psi_c3_batch_t2000      config[3]'s bond dimension AND batch (D=128, 128 clips), 2000 samples
  psi_c4_batch_t4000      config[4]'s bond dimension and per-GPU batch (D=64, 256 clips: more clips than
                          SMs, i.e. the multi-wave single-CTA dispatch), 4000 samples
Only outputs are stored; inputs are regenerated from the recorded seeds (checksums guard them).

    python -m oracle.mint_golden_r2
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cref  # noqa: E402
from oracle.cmps_oracle import HP, PsiCMPSOracle, damped_sine, random_raw_params  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SAMPLE_STRIDE = 125


def sample_noise(hp, L, n, seed):
    """N(0,1) * sigma * sqrt(delta_t), [L, n] float32 (model.py:246 with temp = 1)."""
    z = np.random.default_rng(seed).standard_normal((L, n)).astype(np.float32)
    return z * np.float32(hp.sigma * np.sqrt(hp.delta_t))


def mint_sample(name, D, n, L, seed):
    hp = HP(bond_dim=D, minibatch_size=n)
    raw = random_raw_params(hp, np.random.default_rng(seed))
    o = PsiCMPSOracle(hp, raw, mode="f32", requires_grad=False)
    R, f, p0, A = cref.effective_from_oracle(o)
    noise = sample_noise(hp, L, n, seed + 2)
    t0 = time.time()
    out = cref.psi_sample(R, f, p0, A, hp.sigma, hp.delta_t, noise, mode="f64")
    print(f"C oracle f64 sampler {n}x{L} D={D}: {time.time()-t0:.1f} s; out[0,-1]={out[0,-1]:.6g}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, D=D, n=n, L=L, stride=SAMPLE_STRIDE,
                        sub=out[:, SAMPLE_STRIDE - 1::SAMPLE_STRIDE], last=out[:, -1],
                        rowsum=out.sum(axis=1), absmax=np.abs(out).max(axis=1),
                        noise_checksum=np.float64(np.abs(noise.astype(np.float64)).sum()))


def mint_batch(name, D, B, T, seed):
    hp = HP(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(hp, np.random.default_rng(seed))
    data = damped_sine(B, 64000, hp.delta_t, np.random.default_rng(seed + 1))[:, 3000:3000 + T]
    data = np.ascontiguousarray(data)
    o = PsiCMPSOracle(hp, raw, mode="f32", requires_grad=False)
    R, f, p0, A = cref.effective_from_oracle(o)
    t0 = time.time()
    loss, gR, gf, gp, gA = cref.psi_loss_grad(R, f, p0, A, hp.sigma, hp.delta_t, data, mode="f64")
    print(f"C oracle f64 {B}x{T} D={D}: {time.time()-t0:.1f} s; loss[0:3]={loss[:3]}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, D=D, B=B, T=T, offset=3000,
                        loss_f64=loss, geff_R=gR, geff_f=gf, geff_psi0=gp, geff_A=gA,
                        data_checksum=np.float64(np.abs(data.astype(np.float64)).sum()))


def main():
    which = sys.argv[1:] or ["c2", "c3", "c4"]
    if "c2" in which:
        mint_sample("psi_c2_sample_full", 32, 256, 64000, 200)
    if "c3" in which:
        mint_batch("psi_c3_batch_t2000", 128, 128, 2000, 201)
    if "c4" in which:
        mint_batch("psi_c4_batch_t4000", 64, 256, 4000, 202)


if __name__ == "__main__":
    main()
