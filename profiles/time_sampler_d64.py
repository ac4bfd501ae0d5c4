import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from audio_mps_b200 import HParams, PsiCMPS
dev = torch.device("cuda", 0)
D, n, L = 64, 256, 16000
hp = HParams(minibatch_size=n, bond_dim=D, delta_t=1 / 16000, sigma=0.0001, h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
m = PsiCMPS(hp, device=dev, seed=0)
noise = (torch.randn(L, n, generator=torch.Generator().manual_seed(2)) * m.sigma * np.sqrt(m.delta_t)).to(dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w = m.sample_from_noise(noise); e1.record(); torch.cuda.synchronize()
    print("D=64 sampler 256 x 16000:", e0.elapsed_time(e1), "ms")
