"""Shared helpers for the parity tests (oracle = checker only)."""
import numpy as np
import torch

from oracle.cmps_oracle import HP, PsiCMPSOracle, RhoCMPSOracle, damped_sine, random_raw_params
from audio_mps_b200 import HParams, PsiCMPS, RhoCMPS


def hp_pair(**kw):
    """(oracle HP, product HParams) with identical values."""
    base = dict(minibatch_size=8, bond_dim=8, delta_t=1 / 16000, sigma=0.0001,
                h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100.,
                learning_rate=0.001)
    base.update(kw)
    return HP(**base), HParams(**base)


def set_raw(model, raw):
    """Copy oracle-style raw variables into a product model."""
    with torch.no_grad():
        for k, v in raw.items():
            name = "freqs_raw" if k == "freqs" else k
            getattr(model, name).copy_(torch.as_tensor(np.asarray(v, np.float32)))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def relc(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
