"""Diagnostic: error of the parallel-in-time scan against the float64 oracle as the number of time chunks
changes (the case tests/test_gpu_scan.py::test_scan_random_shapes flagged: D=8, B=2, T=1500)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import hp_pair, set_raw  # noqa: E402
from audio_mps_b200 import PsiCMPS  # noqa: E402
from oracle.cmps_oracle import PsiCMPSOracle, damped_sine, random_raw_params  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(77)
cases = [(3, 1, 97), (8, 2, 1500)]
for D, B, T in cases:
    s1, s2 = int(rng.integers(1 << 30)), int(rng.integers(1 << 30))
ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
raw = random_raw_params(ohp, np.random.default_rng(s1))
data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(s2))
ref = PsiCMPSOracle(ohp, raw, mode="f64").loss_per_clip(data).detach().numpy()
for rep in (1, 2, 4, 8, 16, 37):
    x = np.tile(data, (rep, 1))
    _, php2 = hp_pair(bond_dim=D, minibatch_size=B * rep)
    m = PsiCMPS(php2, device=dev)
    set_raw(m, raw)
    with torch.no_grad():
        ls = m.loss_per_clip(x, time_parallel=True).cpu().numpy()[:B]
        lq = m.loss_per_clip(x, time_parallel=False).cpu().numpy()[:B]
    nvc = max(1, 148 // (B * rep))
    print(f"B={B*rep:3d} nvc~{nvc:3d}: scan err {np.abs(ls-ref)/np.abs(ref)}, seq err {np.abs(lq-ref)/np.abs(ref)}", flush=True)
