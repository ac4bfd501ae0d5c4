"""C2 sampler (D=32, 256 samples x L) for ncu captures. usage: python profiles/prof_sampler.py [L]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS  # noqa: E402

dev = torch.device("cuda", 0)
D, n, L = 32, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 16000
hp = HParams(minibatch_size=n, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
             h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
m = PsiCMPS(hp, device=dev, seed=0)
noise = (torch.randn(L, n, generator=torch.Generator().manual_seed(2)) * m.sigma * np.sqrt(m.delta_t)).to(dev)
for _ in range(2):
    w = m.sample_from_noise(noise)
torch.cuda.synchronize()
print("sampled", tuple(w.shape), float(w.abs().max()))
