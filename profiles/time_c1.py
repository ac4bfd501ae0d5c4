"""C1 / C0 loss+gradient kernel times (CUDA events around the scan kernels), for A/B runs.
usage: python profiles/time_c1.py [D B T]..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, _lib, damped_sine  # noqa: E402

dev = torch.device("cuda", 0)
args = [int(a) for a in sys.argv[1:]] or [32, 64, 64000, 8, 8, 16000]
for D, B, T in zip(args[0::3], args[1::3], args[2::3]):
    hp = HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
                 h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
    m = PsiCMPS(hp, device=dev, seed=0)
    x = torch.from_numpy(damped_sine(B, T, 1 / 16000, np.random.default_rng(1))).to(dev)
    _lib.set_profiling(0, True)
    f, b, s = [], [], []
    for it in range(6):
        m.zero_grad()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.loss_fn(x).backward()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            f.append(_lib.kernel_ms(0, 0)); b.append(_lib.kernel_ms(0, 1)); s.append(e0.elapsed_time(e1))
    print(f"D={D} B={B} T={T}: fwd {min(f):.3f} ms ({min(f)*1e-3/(T-1)*1.965e9:.0f} cyc/step), "
          f"bwd {min(b):.3f} ms ({min(b)*1e-3/(T-1)*1.965e9:.0f} cyc/step), step {min(s):.3f} ms", flush=True)
