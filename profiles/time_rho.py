"""RhoCMPS timings (density-matrix variant, model.py:59-203): loss + gradient and sampling."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, RhoCMPS, damped_sine  # noqa: E402

dev = torch.device("cuda", 0)


def timed(fn, reps=3):
    out = []
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return min(out[1:])


for D, B, T in ((8, 8, 16000), (16, 8, 16000), (32, 8, 16000), (8, 64, 16000), (32, 64, 4000)):
    hp = HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
                 h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
    m = RhoCMPS(hp, device=dev, seed=0)
    x = torch.from_numpy(damped_sine(B, T, hp.delta_t, np.random.default_rng(1))).to(dev)
    with torch.no_grad():
        tf = timed(lambda: m.loss_per_clip(x))

    def step():
        m.zero_grad()
        m.loss_fn(x).backward()
    ts = timed(step)
    noise = (torch.randn(T, B, generator=torch.Generator().manual_seed(2)) * m.sigma * np.sqrt(m.delta_t)).to(dev)
    tsm = timed(lambda: m.sample_from_noise(noise))
    print(f"Rho D={D} B={B} T={T}: fwd {tf:.2f} ms ({tf*1e-3/T*1.965e9:.0f} cyc/step) | fwd+bwd {ts:.2f} ms -> "
          f"{B*T/ts*1e3:.3e} samples/s | sampler {tsm:.2f} ms")
