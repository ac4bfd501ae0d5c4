"""Timings of the other BASELINE configs (kernel + API level, CUDA events).
usage: python profiles/time_configs.py [c2] [c3] [c4] [c0]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, _lib, damped_sine  # noqa: E402

dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["c2", "c4", "c0"]


def hp(D, B):
    return HParams(minibatch_size=B, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
                   h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)


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


if "c2" in which:   # sampling D=32, 256 samples, 4 s clips from a fixed noise tensor
    D, n, L = 32, 256, 64000
    m = PsiCMPS(hp(D, n), device=dev, seed=0)
    noise = (torch.randn(L, n, generator=torch.Generator().manual_seed(2)) * m.sigma * np.sqrt(m.delta_t)).to(dev)
    ms = timed(lambda: m.sample_from_noise(noise))
    print(f"C2 sampler D={D} n={n} L={L}: {ms:.2f} ms  -> {n*L/ms*1e3:.3e} samples/s  ({ms*1e-3/L*1.965e9:.0f} cyc/step)")
if "c4" in which:   # D=64, 256 clips per GPU (global 2048 over 8 GPUs), 4 s clips
    D, B, T = 64, 256, 64000
    m = PsiCMPS(hp(D, B), device=dev, seed=0)
    x = torch.from_numpy(damped_sine(B, T, 1 / 16000, np.random.default_rng(1))).to(dev)
    _lib.set_profiling(0, True)

    def step():
        m.zero_grad()
        m.loss_fn(x).backward()
    ms = timed(step, reps=2)
    print(f"C4 per-GPU D={D} B={B} T={T}: step {ms:.1f} ms (fwd kernel {_lib.kernel_ms(0,0):.1f}, bwd kernel {_lib.kernel_ms(0,1):.1f}) "
          f"-> {B*T/ms*1e3:.3e} samples/s")
if "c0" in which:   # D=8, 8 clips of 1 s
    D, B, T = 8, 8, 16000
    m = PsiCMPS(hp(D, B), device=dev, seed=0)
    x = torch.from_numpy(damped_sine(B, T, 1 / 16000, np.random.default_rng(1))).to(dev)

    def step0():
        m.zero_grad()
        m.loss_fn(x).backward()
    ms = timed(step0)
    print(f"C0 D={D} B={B} T={T}: step {ms:.2f} ms -> {B*T/ms*1e3:.3e} samples/s")
if "c3" in which:   # D=128, 128 clips, 4 s clips: row-split 4-CTA cluster kernels
    D, T = 128, 64000
    for B in (1, 37, 128):
        m = PsiCMPS(hp(D, B), device=dev, seed=0)
        m.time_parallel = "never"
        x = torch.from_numpy(damped_sine(B, T, 1 / 16000, np.random.default_rng(1))).to(dev)
        _lib.set_profiling(0, True)

        def step3():
            m.zero_grad()
            m.loss_fn(x).backward()
        ms = timed(step3, reps=2)
        f, bw = _lib.kernel_ms(0, 0), _lib.kernel_ms(0, 1)
        print(f"C3 D={D} B={B} T={T}: step {ms:.1f} ms (fwd kernel {f:.1f}, bwd kernel {bw:.1f}) "
              f"-> {B*T/ms*1e3:.3e} samples/s; waves {-(-B*4//148)}; "
              f"fwd {f*1e-3/T*1.965e9/max(1,-(-B*4//148)):.0f} cyc/step/wave, bwd {bw*1e-3/T*1.965e9/max(1,-(-B*4//148)):.0f}")
        del m, x
        torch.cuda.empty_cache()
if "big" in which:   # batches beyond one clip per SM pair, D=32 (warp-specialised single-CTA kernels)
    D, T = 32, 16000
    for B in (64, 74, 148, 296, 592, 1184):
        m = PsiCMPS(hp(D, B), device=dev, seed=0)
        x = torch.from_numpy(damped_sine(B, T, 1 / 16000, np.random.default_rng(1))).to(dev)
        _lib.set_profiling(0, True)

        def stepb():
            m.zero_grad()
            m.loss_fn(x).backward()
        ms = timed(stepb, reps=2)
        print(f"big D={D} B={B} T={T}: step {ms:.1f} ms (fwd kernel {_lib.kernel_ms(0,0):.1f}, bwd kernel {_lib.kernel_ms(0,1):.1f}) "
              f"-> {B*T/ms*1e3:.3e} samples/s")
        del m, x
        torch.cuda.empty_cache()
