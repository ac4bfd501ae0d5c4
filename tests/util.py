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


def _rows(a, b, floor):
    """worst row (first axis) of max|a_r - b_r| / max(max|b_r|, floor * max|b|)"""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-300)
    if a.ndim < 2:
        return float(np.abs(a - b).max() / scale) if a.size else 0.0
    ar = a.reshape(a.shape[0], -1)
    br = b.reshape(b.shape[0], -1)
    den = np.maximum(np.abs(br).max(axis=1), floor * scale)
    return float((np.abs(ar - br).max(axis=1) / den).max())


def rel(a, b, floor=1e-3):
    """Relative error of a real array against its reference.  Arrays with >= 2 axes (gradient
    matrices, sample paths [n, L], trajectories [B, ...]) are judged ROW BY ROW (first axis), so a
    small row is not hidden behind a large one; vectors and scalars by their max-norm."""
    return _rows(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), floor)


def relc(a, b, floor=1e-3):
    return _rows(np.asarray(a), np.asarray(b), floor)


def rel_clip(a, b, floor=1e-2):
    """PER-CLIP relative error (north_star: "per-clip log-likelihood within 1e-4 relative"):
    max_b |a_b - r_b| / max(|r_b|, floor * max|r|).  Returns (worst error, worst clip)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    if a.size == 0:
        return 0.0
    den = np.maximum(np.abs(b), floor * max(float(np.abs(b).max()), 1e-300))
    e = np.abs(a - b) / den
    return float(e.max())


def worst_clip(a, b, floor=1e-2):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    den = np.maximum(np.abs(b), floor * max(float(np.abs(b).max()), 1e-300))
    e = np.abs(a - b) / den
    i = int(e.argmax())
    return i, float(e[i]), float(a[i]), float(b[i])


def rel_clip_cond(a, b, abs_terms, kappa=0.1, floor=1e-2):
    """Per-clip error with a condition-aware denominator: max(|r_b|, floor * max|r|, kappa * sum_k |term_kb|).
    Used for the parallel-in-time scan, whose chunk start states carry ~1e-6 of float32 / tf32-split rounding:
    on a clip whose log-terms cancel to a few percent of their absolute sum that is 1e-4 of the (small) loss
    although every term is accurate to 1e-6 (profiles/diag_scan_err.py)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    s = np.asarray(abs_terms, dtype=np.float64).reshape(-1)
    den = np.maximum(np.maximum(np.abs(b), floor * max(float(np.abs(b).max()), 1e-300)), kappa * s)
    return float((np.abs(a - b) / den).max())
