"""Diagnostic: error of the parallel-in-time scan and of the sequential kernels against the float64 oracle,
for the case tests/test_gpu_scan.py::test_scan_random_shapes flagged (D=8, B=2, T=1500) and longer clips."""
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
for D, B, T in [(3, 1, 97), (8, 2, 1500)]:
    s1, s2 = int(rng.integers(1 << 30)), int(rng.integers(1 << 30))
for D, T in [(8, 1500), (8, 3000), (8, 6000), (8, 12000), (32, 1500), (32, 6000), (64, 1500), (64, 6000)]:
    B = 2
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(ohp, np.random.default_rng(s1))
    data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(s2))
    ref = PsiCMPSOracle(ohp, raw, mode="f64").loss_per_clip(data).detach().numpy()
    m = PsiCMPS(php, device=dev)
    set_raw(m, raw)
    with torch.no_grad():
        ls = m.loss_per_clip(data, time_parallel=True).cpu().numpy()
        lq = m.loss_per_clip(data, time_parallel=False).cpu().numpy()
    print(f"D={D} T={T}: loss {ref}, scan err {np.abs(ls-ref)/np.abs(ref)}, seq err {np.abs(lq-ref)/np.abs(ref)}", flush=True)
