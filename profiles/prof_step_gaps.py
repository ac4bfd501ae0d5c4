"""Where the non-kernel time of a C1 training step goes: CUDA-event marks between the phases of
Trainer.step (device timeline) next to the library's own kernel timings."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, _lib, damped_sine  # noqa: E402
from audio_mps_b200.train import Trainer, regulariser  # noqa: E402

dev = torch.device("cuda", 0)
D, B, T = 32, 64, 64000
hp = HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
             h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
m = PsiCMPS(hp, device=dev, seed=0)
tr = Trainer(m)
x = torch.from_numpy(damped_sine(B, T, hp.delta_t, np.random.default_rng(1))).to(dev)
_lib.set_profiling(0, True)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for it in range(6):
    torch.cuda.synchronize()
    e0 = ev()
    lpc = m.loss_per_clip(x)
    e1 = ev()
    obj = lpc.sum() / B + regulariser(m)
    tr.opt.zero_grad(set_to_none=True)
    e2 = ev()
    obj.backward()
    e3 = ev()
    tr.opt.step()
    e4 = ev()
    torch.cuda.synchronize()
    f, bw = _lib.kernel_ms(0, 0), _lib.kernel_ms(0, 1)
    if it >= 2:
        print(f"step {e0.elapsed_time(e4):.2f} ms | loss_per_clip {e0.elapsed_time(e1):.2f} (fwd kernel {f:.2f}) | "
              f"objective {e1.elapsed_time(e2):.2f} | backward {e2.elapsed_time(e3):.2f} (bwd kernel {bw:.2f}) | "
              f"adam {e3.elapsed_time(e4):.2f}")
