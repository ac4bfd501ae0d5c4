"""One parallel-in-time scan forward (D=64, B=1) for ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, damped_sine  # noqa: E402

dev = torch.device("cuda", 0)
D, B, T = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 64000
hp = HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
             h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
m = PsiCMPS(hp, device=dev, seed=0)
x = torch.from_numpy(damped_sine(B, T, hp.delta_t, np.random.default_rng(1))).to(dev)
for _ in range(2):
    l = m.loss_per_clip_scan(x)
torch.cuda.synchronize()
print("scan loss", l.detach().cpu().numpy())
