"""ctypes wrapper of oracle/libcmps_ref.so (C restatement of the reference scan).
TEST INFRASTRUCTURE ONLY -- see cmps_ref.c."""
import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcmps_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _c64_ri(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.complex64))
    return a.view(np.float32)


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def psi_loss(R, freqs, psi0, A, sigma, delta_t, x, mode="f64"):
    lib = load()
    x = _f32(x)
    B, T = x.shape
    D = len(freqs)
    Rr, fr, pr = _c64_ri(R), _f32(freqs), _c64_ri(psi0)
    loss = np.zeros(B, np.float64)
    fn = getattr(lib, f"cmps_psi_loss_{mode}")
    fn(C.c_int(D), C.c_int(B), C.c_int(T), _p(Rr, C.c_float), _p(fr, C.c_float), _p(pr, C.c_float),
       C.c_float(A), C.c_float(sigma), C.c_double(delta_t), _p(x, C.c_float), _p(loss, C.c_double))
    return loss


def psi_loss_grad(R, freqs, psi0, A, sigma, delta_t, x, w=None, mode="f64"):
    """(loss[B], gR complex[D,D], gf[D], gpsi0 complex[D], gA) wrt the EFFECTIVE parameters."""
    lib = load()
    x = _f32(x)
    B, T = x.shape
    D = len(freqs)
    Rr, fr, pr = _c64_ri(R), _f32(freqs), _c64_ri(psi0)
    w = np.ascontiguousarray(np.full(B, 1.0 / B) if w is None else np.asarray(w, np.float64))
    loss = np.zeros(B, np.float64)
    gR = np.zeros(2 * D * D, np.float64)
    gf = np.zeros(D, np.float64)
    gp = np.zeros(2 * D, np.float64)
    gA = np.zeros(1, np.float64)
    fn = getattr(lib, f"cmps_psi_loss_grad_{mode}")
    fn(C.c_int(D), C.c_int(B), C.c_int(T), _p(Rr, C.c_float), _p(fr, C.c_float), _p(pr, C.c_float),
       C.c_float(A), C.c_float(sigma), C.c_double(delta_t), _p(x, C.c_float), _p(w, C.c_double),
       _p(loss, C.c_double), _p(gR, C.c_double), _p(gf, C.c_double), _p(gp, C.c_double), _p(gA, C.c_double))
    gRc = gR.reshape(D, D, 2) @ np.array([1, 1j])
    gpc = gp.reshape(D, 2) @ np.array([1, 1j])
    return loss, gRc, gf, gpc, float(gA[0])


def psi_sample(R, freqs, psi0, A, sigma, delta_t, noise, mode="f64"):
    lib = load()
    noise = _f32(noise)
    L, n = noise.shape
    D = len(freqs)
    Rr, fr, pr = _c64_ri(R), _f32(freqs), _c64_ri(psi0)
    out = np.zeros((n, L), np.float64)
    fn = getattr(lib, f"cmps_psi_sample_{mode}")
    fn(C.c_int(D), C.c_int(n), C.c_int(L), _p(Rr, C.c_float), _p(fr, C.c_float), _p(pr, C.c_float),
       C.c_float(A), C.c_float(sigma), C.c_double(delta_t), _p(noise, C.c_float), _p(out, C.c_double))
    return out


def effective_from_oracle(o):
    """(R, freqs, psi0, A) numpy arrays of a PsiCMPSOracle."""
    return (o.R.detach().numpy().astype(np.complex64), o.freqs.detach().numpy().astype(np.float32),
            o.psi_0.detach().numpy().astype(np.complex64), float(o.A.detach()))


def bench_loss_grad(D, B, t_sample, hp_kw, mode="f32"):
    """samples/s of the float32 C/OpenMP port, loss + adjoint gradient, on B clips x t_sample steps."""
    from oracle.cmps_oracle import HP, PsiCMPSOracle, damped_sine, random_raw_params
    hp = HP(**hp_kw)
    raw = random_raw_params(hp, np.random.default_rng(0))
    o = PsiCMPSOracle(hp, raw, mode="f32", requires_grad=False)
    R, f, p0, A = effective_from_oracle(o)
    full = damped_sine(B, 64000, hp.delta_t, np.random.default_rng(1))
    x = np.ascontiguousarray(full[:, 4000:4000 + t_sample + 1])
    psi_loss_grad(R, f, p0, A, hp.sigma, hp.delta_t, x[:, :50], mode=mode)
    t0 = time.perf_counter()
    psi_loss_grad(R, f, p0, A, hp.sigma, hp.delta_t, x, mode=mode)
    return B * t_sample / (time.perf_counter() - t0)
