"""Parallel-in-time tensor-core scan vs the sequential kernel at small batch (CUDA events)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, _lib, damped_sine  # noqa: E402

dev = torch.device("cuda", 0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
for D, B in ((64, 1), (64, 2), (64, 4), (64, 8), (64, 16), (64, 32), (32, 1), (32, 2), (32, 4), (32, 8), (16, 1), (16, 4), (8, 1), (8, 2)):
    hp = HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
                 h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
    m = PsiCMPS(hp, device=dev, seed=0)
    x = torch.from_numpy(damped_sine(B, T, hp.delta_t, np.random.default_rng(1))).to(dev)

    def timed(fn, reps=3):
        out = []
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        return min(out[1:]), r
    with torch.no_grad():
        ts, ls = timed(lambda: m.loss_per_clip(x, time_parallel=False))
        tp, lp = timed(lambda: m.loss_per_clip_scan(x))
    err = float(((lp - ls).abs() / ls.abs()).max())
    print(f"D={D} B={B} T={T}: sequential {ts:.2f} ms | tcgen05 scan {tp:.2f} ms | speed-up {ts/tp:.1f}x | "
          f"rel diff {err:.1e} | scan {B*T/tp*1e3:.3e} samples/s")
    # loss + gradient (forward with saved trajectories, then the adjoint)
    ps = list(m.parameters())
    tsg, gs = timed(lambda: torch.autograd.grad(m.loss_per_clip(x, time_parallel=False).mean(), ps))
    tpg, gp = timed(lambda: torch.autograd.grad(m.loss_per_clip_scan(x).mean(), ps))
    gerr = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(gp, gs))
    print(f"    fwd+bwd: sequential {tsg:.2f} ms | scan {tpg:.2f} ms | speed-up {tsg/tpg:.1f}x | "
          f"grad rel diff {gerr:.1e} | scan {B*T/tpg*1e3:.3e} samples/s")
