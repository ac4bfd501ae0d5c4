import os

import numpy as np

from oracle.cmps_oracle import HP, damped_sine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))


def psi_case(name):
    """(HP, raw dict, data, golden dict) of a minted Psi case (oracle/mint_golden.py:psi_case)."""
    g = load(name)
    D, delta_t, sigma, h_reg, r_reg, A = g["hp"]
    B, T, seed = int(g["B"]), int(g["T"]), int(g["seed"])
    hp = HP(minibatch_size=B, bond_dim=int(D), delta_t=float(delta_t), sigma=float(sigma),
            h_reg=float(h_reg), r_reg=float(r_reg), A=float(A))
    raw = {k[4:]: g[k] for k in g if k.startswith("raw_")}
    data = g["data"] if "data" in g else damped_sine(B, T, hp.delta_t, np.random.default_rng(seed + 1))
    return hp, raw, data, g


PSI_CASES = ["psi_testhp_d7", "psi_train_d8_t1500", "psi_c0_d8_t16000", "psi_d32_t600"]
