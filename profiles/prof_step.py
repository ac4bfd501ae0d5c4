"""One PsiCMPS training step (fwd scan + adjoint bwd) for ncu captures.
usage: python profiles/prof_step.py [D] [B] [T] [reps] [K]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, damped_sine  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
K = int(sys.argv[5]) if len(sys.argv) > 5 else None
dev = torch.device("cuda", 0)
hp = HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
             h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
model = PsiCMPS(hp, device=dev, seed=0)
if K is not None:
    model.checkpoint_every = K
x = torch.from_numpy(damped_sine(B, T, hp.delta_t, np.random.default_rng(1))).to(dev)
for _ in range(reps):
    model.zero_grad()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    loss = model.loss_fn(x)
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    print(f"D={D} B={B} T={T} loss={float(loss):.6f} fwd {e0.elapsed_time(e1):.3f} ms bwd {e1.elapsed_time(e2):.3f} ms")
