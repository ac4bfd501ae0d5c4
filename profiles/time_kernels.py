"""Kernel-only timings (CUDA events around the fwd / bwd scan kernels inside the library).
usage: python profiles/time_kernels.py [D] [B] [T] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, _lib, damped_sine  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda", 0)
hp = HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
             h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
model = PsiCMPS(hp, device=dev, seed=0)
x = torch.from_numpy(damped_sine(B, T, hp.delta_t, np.random.default_rng(1))).to(dev)
_lib.set_profiling(0, True)
f, b = [], []
for r in range(reps):
    model.zero_grad()
    loss = model.loss_fn(x)
    loss.backward()
    torch.cuda.synchronize()
    if r > 0:
        f.append(_lib.kernel_ms(0, 0))
        b.append(_lib.kernel_ms(0, 1))
steps = T - 1
clk = 1.965e9
print(f"D={D} B={B} T={T} cluster={'off' if os.environ.get('AMPS_NO_CLUSTER') == '1' else 'auto'} "
      f"loss={float(loss.detach()):.6f}  fwd {min(f):.3f} ms ({min(f)*1e-3/steps*clk:.0f} cyc/step)  "
      f"bwd {min(b):.3f} ms ({min(b)*1e-3/steps*clk:.0f} cyc/step)")
