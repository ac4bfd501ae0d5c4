"""Checkpoint interval K: loss+gradient time and workspace, CUDA events.
usage: python profiles/time_ckpt.py [c1] [c4] [c3] [b2048]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import HParams, PsiCMPS, _lib, damped_sine  # noqa: E402

dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["c1"]
CFG = {"c1": (32, 64, 64000, (1, 512, 1024, 2048, 4096, 8192)), "c4": (64, 256, 64000, (1, 2048, 4096)),
       "c3": (128, 128, 64000, (1, 2048)), "c4k1": (64, 256, 64000, (1,)), "c3k1": (128, 128, 64000, (1,)), "b2048": (64, 2048, 64000, (2048,)), "c0": (8, 8, 16000, (1, 1024, 2048)),
       "b16": (32, 16, 64000, (1, 2048)), "b32": (32, 32, 64000, (1, 2048))}


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


lib = _lib.load()
for name in which:
    D, B, T, Ks = CFG[name]
    m = PsiCMPS(hp(D, B), device=dev, seed=0)
    x = torch.from_numpy(damped_sine(min(B, 256), T, 1 / 16000, np.random.default_rng(1))).to(dev)
    if B > x.shape[0]:
        x = x.repeat(B // x.shape[0], 1).contiguous()
    _lib.set_profiling(0, True)
    g1 = None
    for K in Ks:
        m.checkpoint_every = K

        def step():
            m.zero_grad()
            m.loss_fn(x).backward()
        ms = timed(step, reps=2 if B * D >= 8192 else 3)
        _lib.set_profiling(0, False)      # without the per-kernel events partial waves are pipelined (launch_waves)
        ms_pipe = timed(step, reps=2 if B * D >= 8192 else 3)
        _lib.set_profiling(0, True)
        g = m.Rx.grad.detach().clone()
        if g1 is None:
            g1 = g
        dg = float((g - g1).abs().max() / g1.abs().max())
        tiles = ""
        try:
            tiles = f", tiles {_lib.kernel_ms(0,3):.2f}, sx {_lib.kernel_ms(0,2):.2f}"
        except Exception:
            pass
        print(f"{name} D={D} B={B} T={T} K={K}: step {ms:.2f} ms (fwd {_lib.kernel_ms(0,0):.2f}, bwd {_lib.kernel_ms(0,1):.2f}{tiles}) "
              f"; step without kernel events {ms_pipe:.2f} ms -> {B*T/ms_pipe*1e3:.3e} samples/s; workspace {lib.amps_psi_workspace_bytes_k(D,B,T,K)/1e6:.1f} MB; "
              f"dRx vs first K {dg:.1e}", flush=True)
