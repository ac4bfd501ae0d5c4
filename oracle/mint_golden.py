"""Mint the golden vectors under tests/golden/ from the op-for-op oracle (oracle/cmps_oracle.py).

    python -m oracle.mint_golden            (from the repo root; ~1 minute)

The reference itself cannot run in this image (TensorFlow 1.x is not installable), and its
own tests hold no numbers, so these fixtures ARE the pin: they freeze the restatement's outputs so a
later edit of the oracle, the C port or the kernels shows up as a diff.  Inputs are stored with the
outputs (small cases) or regenerated from the recorded seeds (C0).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cmps_oracle import (HP, PsiCMPSOracle, RhoCMPSOracle, damped_sine, grads_of,  # noqa: E402
                                random_raw_params, ref_test_hparams, total_loss)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def psi_case(name, hp, B, T, seed, both_modes=True):
    raw = random_raw_params(hp, np.random.default_rng(seed))
    data = damped_sine(B, T, hp.delta_t, np.random.default_rng(seed + 1))
    out = {"seed": seed, "B": B, "T": T, "hp": np.array([hp.bond_dim, hp.delta_t, hp.sigma, hp.h_reg, hp.r_reg, hp.A])}
    out.update({f"raw_{k}": np.asarray(v) for k, v in raw.items()})
    if B * T <= 8 * 2048:
        out["data"] = data
    for mode in (("f32", "f64") if both_modes else ("f64",)):
        o = PsiCMPSOracle(hp, raw, mode=mode)
        lpc = o.loss_per_clip(data)
        g = grads_of(o, total_loss(o, data))
        out[f"loss_{mode}"] = lpc.detach().numpy()
        for k, v in g.items():
            out[f"grad_{k}_{mode}"] = v
        if mode == "f64":
            eff = torch.autograd.grad(PsiCMPSOracle(hp, raw, mode=mode).loss(data), [], allow_unused=True) if False else None
            o2 = PsiCMPSOracle(hp, raw, mode=mode)
            ge = torch.autograd.grad(o2.loss(data), [o2.R, o2.freqs, o2.psi_0, o2.A])
            out["R_eff"] = o2.R.detach().numpy()
            out["freqs_eff"] = o2.freqs.detach().numpy()
            out["psi0"] = o2.psi_0.detach().numpy()
            out["geff_R"], out["geff_f"], out["geff_psi0"], out["geff_A"] = [t.numpy() for t in ge]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: out[k] for k in ("loss_f64",)})


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. the reference's test hparams (tests/test_model.py:13-14): D=7, B=8, T=256
    psi_case("psi_testhp_d7", ref_test_hparams(), 8, 256, seed=10)
    # 2. train.py hparams, small
    psi_case("psi_train_d8_t1500", HP(), 8, 1500, seed=20)
    # 3. BASELINE config[0]: D=8, B=8, 1 s @ 16 kHz (T=16000); data regenerated from the seed
    psi_case("psi_c0_d8_t16000", HP(), 8, 16000, seed=30, both_modes=True)
    # 4. D=32 (the headline bond dimension), short
    psi_case("psi_d32_t600", HP(bond_dim=32, minibatch_size=4), 4, 600, seed=40)

    # 5. the two-level system of tests/test_model.py:140-158 with a fixed noise tensor
    hp = HP(minibatch_size=8, bond_dim=2, delta_t=1 / 16000, sigma=1, initial_rank=None, A=1.,
            h_reg=2 / (np.pi * 16000) ** 2, r_reg=2 / (np.pi * 16000) ** 2)
    R = np.array([[0, 1], [0, 0]], dtype=np.complex64)
    fr = np.array([10, -10], dtype=np.float32)
    psi_xy = {"psi_x": np.array([0.6, 0.3], np.float32), "psi_y": np.array([0.1, -0.7], np.float32)}
    noise = (np.random.default_rng(50).standard_normal((512, 2)) * hp.sigma * np.sqrt(hp.delta_t)).astype(np.float32)
    q = PsiCMPSOracle(hp, psi_xy, R_in=R, freqs_in=fr, mode="f64")
    q32 = PsiCMPSOracle(hp, psi_xy, R_in=R, freqs_in=fr, mode="f32")
    W = np.array([[0.8, 0.1 + 0.2j], [0.3j, 0.5]], dtype=np.complex64)
    r = RhoCMPSOracle(hp, None, W_in=W, R_in=R, freqs_in=fr, mode="f64")
    np.savez_compressed(os.path.join(OUT, "qubit_sampling.npz"), noise=noise, psi_x=psi_xy["psi_x"],
                        psi_y=psi_xy["psi_y"], W=W,
                        psi_sample_f64=q.sample_from_noise(noise).detach().numpy(),
                        psi_sample_f32=q32.sample_from_noise(noise).detach().numpy(),
                        rho_sample_f64=r.sample_from_noise(noise).detach().numpy(),
                        rho_purity_f64=r.purity_from_noise(noise).detach().numpy())

    # 6. sampler at D=7 / D=32 from fixed noise
    for D, n, L, seed in ((7, 5, 256, 60), (32, 3, 400, 61)):
        hp = HP(bond_dim=D)
        raw = random_raw_params(hp, np.random.default_rng(seed))
        noise = (np.random.default_rng(seed + 1).standard_normal((L, n)) * hp.sigma * np.sqrt(hp.delta_t)).astype(np.float32)
        o = PsiCMPSOracle(hp, raw, mode="f64")
        np.savez_compressed(os.path.join(OUT, f"psi_sample_d{D}.npz"), noise=noise, seed=seed,
                            **{f"raw_{k}": np.asarray(v) for k, v in raw.items()},
                            sample_f64=o.sample_from_noise(noise).detach().numpy())

    # 7. rho, test hparams: loss + trace-normalised trajectory checksum
    hp = ref_test_hparams()
    raw = random_raw_params(hp, np.random.default_rng(70), rho=True)
    data = damped_sine(8, 256, hp.delta_t, np.random.default_rng(71))
    o = RhoCMPSOracle(hp, raw, mode="f64")
    tr = o.rho_evolve_with_data(data).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "rho_testhp_d7.npz"), data=data, seed=70,
                        **{f"raw_{k}": np.asarray(v) for k, v in raw.items()},
                        loss_f64=o.loss_per_clip(data).detach().numpy(),
                        loss_f32=RhoCMPSOracle(hp, raw, mode="f32").loss_per_clip(data).detach().numpy(),
                        traj_last=tr[:, -1], traj_abs_sum=np.abs(tr).sum(axis=(2, 3)))


if __name__ == "__main__":
    main()
