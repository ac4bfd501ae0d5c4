"""Diagnostic: gradient error of the sequential (tensor-core tiles or AMPS_NO_TC_TILES=1) and scan paths vs the
float64 oracle for the last case of tests/test_gpu_scan.py::test_scan_random_shapes (D=64, B=2, T=4097)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import hp_pair, rel, set_raw  # noqa: E402
from audio_mps_b200 import PsiCMPS  # noqa: E402
from oracle.cmps_oracle import PsiCMPSOracle, damped_sine, grads_of, random_raw_params  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(77)
cases = [(3, 1, 97), (8, 2, 1500), (17, 3, 640), (32, 5, 2049), (33, 7, 333), (64, 16, 700),
         (12, 20, 130), (64, 37, 65), (16, 74, 100), (9, 75, 64), (5, 149, 40), (64, 2, 4097)]
for D, B, T in cases:
    s1, s2 = int(rng.integers(1 << 30)), int(rng.integers(1 << 30))
NAMES = ("Rx", "Ry", "freqs_raw", "psi_x", "psi_y", "A")
ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
raw = random_raw_params(ohp, np.random.default_rng(s1))
data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(s2))
w = np.linspace(0.7, 1.3, B) / B
o = PsiCMPSOracle(ohp, raw, mode="f64")
gref = grads_of(o, (o.loss_per_clip(data) * torch.as_tensor(w)).sum())
m = PsiCMPS(php, device=dev)
set_raw(m, raw)
ps = [getattr(m, n) for n in NAMES]
wt = torch.as_tensor(w, dtype=torch.float32, device=dev)
for tp in (False, True):
    g = torch.autograd.grad((m.loss_per_clip(data, time_parallel=tp) * wt).sum(), ps)
    print("scan" if tp else "seq ", {n: f"{rel(a.cpu().numpy(), gref['freqs' if n == 'freqs_raw' else n]):.2e}" for n, a in zip(NAMES, g)}, flush=True)
