"""Mint the FULL-LENGTH goldens (BASELINE config C1 at full size: D=32, 64 clips x 64000 samples; and 4
full-length clips at the bond dimensions of C4 (64) and C3 (128)) with the C/OpenMP
restatement in float64 (oracle/cmps_ref.c: lab frame, per-step normalisation, SURVEY App. B adjoint --
an implementation independent of the CUDA chain form).  Takes a few minutes of CPU; only the
outputs are stored (per-clip losses and the packed effective-parameter gradient), the inputs are
regenerated from the recorded seeds.

    python -m oracle.mint_golden_c1
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cref  # noqa: E402
from oracle.cmps_oracle import HP, PsiCMPSOracle, damped_sine, random_raw_params  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def mint(name, D, B, T, seed):
    hp = HP(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(hp, np.random.default_rng(seed))
    data = damped_sine(B, T, hp.delta_t, np.random.default_rng(seed + 1))
    o = PsiCMPSOracle(hp, raw, mode="f32", requires_grad=False)     # float32 effective parameters, as the kernels get
    R, f, p0, A = cref.effective_from_oracle(o)
    t0 = time.time()
    loss, gR, gf, gp, gA = cref.psi_loss_grad(R, f, p0, A, hp.sigma, hp.delta_t, data, mode="f64")
    print(f"C oracle f64, {B}x{T} D={D}: {time.time()-t0:.1f} s; loss[0:4] = {loss[:4]}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, D=D, B=B, T=T,
                        R_eff=R, freqs_eff=f, psi0=p0, A=A, loss_f64=loss,
                        geff_R=gR, geff_f=gf, geff_psi0=gp, geff_A=gA,
                        data_checksum=np.float64(np.abs(data.astype(np.float64)).sum()))


def main():
    mint("psi_c1_full", 32, 64, 64000, 100)      # BASELINE config[1], full size
    mint("psi_c4_d64_full_length", 64, 4, 64000, 101)    # config[4]'s bond dimension, 4 full-length clips
    mint("psi_c3_d128_full_length", 128, 4, 64000, 102)  # config[3]'s bond dimension, 4 full-length clips


if __name__ == "__main__":
    main()
