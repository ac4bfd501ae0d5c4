"""Randomised parity sweep on a GPU: loss and gradients of PsiCMPS against the float64 oracle over random
(D, B, T, K) -- every kernel family, ragged chunks / windows / time splits, batches around the wave boundaries
(D = 64: 148 / 296 clips; D > 64: 32..74 clusters).  usage: python profiles/fuzz_parity.py [n_cases] [seed] [tmax_big]
Prints one line per case and a summary; exit code 1 if a case is out of tolerance (loss 1e-4 per clip with the
conditioning-aware bound of tests/util.rel_clip_cond, gradients 1e-3 per row)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_mps_b200 import PsiCMPS  # noqa: E402
from oracle.cmps_oracle import PsiCMPSOracle, damped_sine, grads_of, random_raw_params  # noqa: E402
from tests.util import hp_pair, rel, rel_clip_cond, set_raw  # noqa: E402

NAMES = ("Rx", "Ry", "freqs_raw", "psi_x", "psi_y", "A")


def run(n_cases=40, seed=0, tmax_big=60, verbose=True):
    """Returns the list of failing cases (empty = all within tolerance)."""
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda", 0)
    failures = []
    for case in range(n_cases):
        err = one_case(case, seed, rng, dev, tmax_big, verbose)
        if err is not None:
            failures.append(err)
    return failures


def one_case(case, seed, rng, dev, tmax_big, verbose):
    if True:
        fam = rng.integers(0, 5)
        if fam == 0:
            D, B = int(rng.integers(1, 33)), int(rng.integers(1, 12))
        elif fam == 1:
            D, B = int(rng.integers(2, 33)), int(rng.integers(70, 160))          # single-CTA family (2 B > #SMs)
        elif fam == 2:
            D, B = int(rng.integers(33, 65)), int(rng.choice([1, 3, 7, 147, 149, 160, 297, 300]))
        elif fam == 3:
            D, B = int(rng.integers(65, 129)), int(rng.choice([1, 2, 5, 31, 33, 34, 40, 66, 75, 80]))
        else:
            D, B = int(rng.choice([8, 32, 64, 128])), int(rng.integers(1, 6))
        big = B > 20
        T = int(rng.integers(2, tmax_big)) if big else int(rng.integers(2, 1500))
        if D > 64 and not big:
            T = min(T, 400)
        K = rng.choice([None, None, 1, 16, 33, 200, 2048])
        ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
        raw = random_raw_params(ohp, np.random.default_rng(1000 * (seed + 1) + case))
        data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(5000 * (seed + 1) + case))
        m = PsiCMPS(php, device=dev)
        set_raw(m, raw)
        if K is not None:
            m.checkpoint_every = int(K)
        w = torch.as_tensor(rng.uniform(0.5, 1.5, B).astype(np.float32), device=dev) / B
        lpc = m.loss_per_clip(data)
        (lpc * w).sum().backward()
        torch.cuda.synchronize()
        o = PsiCMPSOracle(ohp, raw, mode="f64")
        ref = o.loss_per_clip(data)
        _, absterms = o.loss_and_abs_terms(data)
        gref = grads_of(o, (ref * torch.as_tensor(w.cpu().numpy(), dtype=torch.float64)).sum())
        el = rel_clip_cond(lpc.detach().cpu().numpy(), ref.detach().numpy(), absterms)
        eg = max(rel(getattr(m, n).grad.cpu().numpy(), gref["freqs" if n == "freqs_raw" else n]) for n in NAMES)
        ok = el <= 1e-4 and eg <= 1e-3 and np.isfinite(el) and np.isfinite(eg)
        if verbose:
            print(f"case {case:3d} D={D:3d} B={B:3d} T={T:4d} K={K}: loss {el:.2e} grad {eg:.2e} {'ok' if ok else 'FAIL'}", flush=True)
        return None if ok else (case, D, B, T, K, el, eg)


def run_samplers(n_cases=20, seed=0, verbose=True):
    """PsiCMPS.sample_from_noise / RhoCMPS.sample_from_noise, trajectories and the Rho loss + gradient against the
    oracle over random shapes (samples 1e-3 per waveform given the same noise; Rho gradients 1e-3)."""
    from audio_mps_b200 import RhoCMPS
    from oracle.cmps_oracle import RhoCMPSOracle
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda", 0)
    failures = []
    for case in range(n_cases):
        rho = case % 3 == 2
        D = int(rng.integers(1, 33)) if rho else int(rng.choice([rng.integers(1, 33), rng.integers(33, 65), rng.integers(65, 129)]))
        n = int(rng.integers(1, 7)) if (rho or D > 64) else int(rng.choice([1, 3, 9, 150, 300]))
        L = int(rng.integers(1, 300 if n < 20 else 40))
        ohp, php = hp_pair(bond_dim=D, minibatch_size=max(n, 1))
        raw = random_raw_params(ohp, np.random.default_rng(7000 * (seed + 1) + case), rho=rho) if rho else \
            random_raw_params(ohp, np.random.default_rng(7000 * (seed + 1) + case))
        noise = (np.random.default_rng(9000 + case).standard_normal((L, n)) * ohp.sigma * np.sqrt(ohp.delta_t)).astype(np.float32)
        if rho:
            o, m = RhoCMPSOracle(ohp, raw, mode="f64"), RhoCMPS(php, device=dev)
        else:
            o, m = PsiCMPSOracle(ohp, raw, mode="f64"), PsiCMPS(php, device=dev)
        set_raw(m, raw)
        es = rel(m.sample_from_noise(noise).cpu().numpy(), o.sample_from_noise(noise).detach().numpy())
        eg = 0.0
        if rho and L >= 2:
            data = damped_sine(n, L + 1, ohp.delta_t, np.random.default_rng(case))
            ref = o.loss_per_clip(data)
            gref = grads_of(o, ref.mean())
            m.loss_per_clip(data).mean().backward()
            eg = max(rel(getattr(m, k).grad.cpu().numpy(), gref["freqs" if k == "freqs_raw" else k])
                     for k in ("Rx", "Ry", "freqs_raw", "Wx", "Wy", "A"))
        ok = es <= 1e-3 and eg <= 1e-3 and np.isfinite(es) and np.isfinite(eg)
        if verbose:
            print(f"sampler case {case:3d} {'rho' if rho else 'psi'} D={D:3d} n={n:3d} L={L:3d}: samples {es:.2e} "
                  f"rho-grad {eg:.2e} {'ok' if ok else 'FAIL'}", flush=True)
        if not ok:
            failures.append((case, rho, D, n, L, es, eg))
    return failures


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tmax_big = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    t0 = time.time()
    failures = run(n_cases, seed, tmax_big)
    failures += run_samplers(max(n_cases // 2, 1), seed)
    print(f"{n_cases - len(failures)}/{n_cases} ok in {time.time() - t0:.0f} s")
    sys.exit(1 if failures else 0)
