"""CPU oracle for the continuous-MPS ("AudioMPS") time-step scan.

TEST INFRASTRUCTURE ONLY.  Nothing under ``audio_mps_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or as the
thing-timed-as-the-reference, never as the product path.

PARITY UNPINNED.  The reference is TensorFlow-1.x graph code
(/root/reference/model.py); TensorFlow cannot be installed in this image and
the reference's own tests (/root/reference/tests/test_model.py) hold structural
properties only -- no golden numbers.  This file is therefore an op-for-op
restatement of model.py in PyTorch-CPU (same op order, same dtypes, the same
float32 ``t += dt`` running sum), and the golden vectors in ``tests/golden`` are
minted from it by ``oracle/mint_golden.py``.  The properties the reference does
test are ported in ``tests/test_oracle_properties.py``.

Two arithmetic modes:

* ``"f32"``  -- float32 / complex64 everywhere, as the reference runs.
* ``"f64"``  -- float64 / complex128 arithmetic, but with the *definition* of
  the function kept identical to the reference: parameters and data are the
  same float32 values, ``t_k`` is the float32 running sum and the phase angle is
  the float32 product ``fl32(f_c * t_k)`` (model.py:16,157,281,304-305).  This
  is the "exact arithmetic" value both the reference and the CUDA path
  approximate; it is what error budgets are measured against.

Every function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


# ----------------------------------------------------------------------------
# hparams (tf.contrib.training.HParams stand-in; train.py:41-43, tests/test_model.py:13-14)
# ----------------------------------------------------------------------------
@dataclass
class HP:
    minibatch_size: int = 8
    bond_dim: int = 8
    delta_t: float = 1.0 / 16000
    sigma: float = 0.0001
    h_reg: float = 200.0 / (np.pi * 16000) ** 2
    r_reg: float = 0.1
    initial_rank: Optional[int] = None
    A: float = 100.0
    learning_rate: float = 0.001


def ref_test_hparams() -> HP:
    """hparams of /root/reference/tests/test_model.py:13-14."""
    return HP(minibatch_size=8, bond_dim=7, delta_t=1 / 16000, sigma=0.0001,
              initial_rank=None, A=100.0,
              h_reg=2 / (np.pi * 16000) ** 2, r_reg=2 / (np.pi * 16000))


def t_table(n: int, delta_t: float) -> np.ndarray:
    """float32 running sum t_0=0, t_{k+1}=fl32(t_k+fl32(delta_t)) (model.py:16,281)."""
    dt32 = np.float32(delta_t)
    inc = np.full(n, dt32, dtype=np.float32)
    inc[0] = 0.0
    # np.add.accumulate on float32 is a strict left-to-right float32 sum
    return np.add.accumulate(inc, dtype=np.float32)


def damped_sine(batch: int, length: int, delta_t: float, rng: np.random.Generator) -> np.ndarray:
    """Synthetic clips of /root/reference/data.py:8-22 (float32 [batch, length]).

    261.6 Hz sine, 0.1 s decay, onset delay ~ Gamma(shape 2, rate 2/(length/100)).
    """
    freq = 261.6
    decay_time = 0.1
    delay_time = length / 100
    delays = rng.gamma(shape=2.0, scale=delay_time / 2.0, size=(batch, 1)).astype(np.float32)
    input_range = np.arange(length, dtype=np.float32)[None, :]
    times = ((input_range - delays) * np.float32(delta_t)).astype(np.float32)
    wave = 0.5 * (np.sign(times) + 1) * np.sin(2 * np.pi * freq * times) * np.exp(-times / decay_time)
    return wave.astype(np.float32)


def random_raw_params(hp: HP, rng: np.random.Generator, rho: bool = False) -> dict:
    """Raw trainables with the reference's initialisers (model.py:36-39,49-50,218-219,125-126)."""
    D = hp.bond_dim
    out = {
        "Rx": rng.standard_normal((D, D)).astype(np.float32),
        "Ry": rng.standard_normal((D, D)).astype(np.float32),
        "freqs": rng.standard_normal(D).astype(np.float32),
        "A": np.float32(hp.A),
    }
    if rho:
        rank = hp.initial_rank if hp.initial_rank is not None else D
        lim = math.sqrt(6.0 / (rank + D))  # glorot-uniform, SURVEY 3.4
        out["Wx"] = rng.uniform(-lim, lim, (rank, D)).astype(np.float32)
        out["Wy"] = rng.uniform(-lim, lim, (rank, D)).astype(np.float32)
    else:
        lim = math.sqrt(3.0 / D)  # glorot-uniform on a [D] vector
        out["psi_x"] = rng.uniform(-lim, lim, D).astype(np.float32)
        out["psi_y"] = rng.uniform(-lim, lim, D).astype(np.float32)
    return out


# ----------------------------------------------------------------------------
# CMPS base: parameterisation (model.py:9-52)
# ----------------------------------------------------------------------------
class CMPSOracle:
    def __init__(self, hp: HP, raw: Optional[dict] = None, freqs_in=None, R_in=None,
                 mode: str = "f32", requires_grad: bool = True):
        assert mode in ("f32", "f64")
        self.mode = mode
        self.rdt = torch.float32 if mode == "f32" else torch.float64
        self.cdt = torch.complex64 if mode == "f32" else torch.complex128
        self.hp = hp
        self.bond_d = hp.bond_dim
        self.h_reg = hp.h_reg
        self.r_reg = hp.r_reg
        self.delta_t = hp.delta_t                     # python double (model.py:15)
        self.dt32 = np.float32(hp.delta_t)            # model.py:16
        self.sigma = hp.sigma                         # python constant (model.py:21)
        raw = dict(raw) if raw is not None else {}

        def leaf(v):
            # raw variables are float32 VALUES in both modes
            t = torch.tensor(np.asarray(v, dtype=np.float32), dtype=self.rdt)
            t.requires_grad_(requires_grad)
            return t

        self.vars = {}
        self.vars["A"] = leaf(raw.get("A", hp.A))     # model.py:19
        self.A = self.vars["A"]

        if R_in is not None:                          # model.py:31-33 (no rsqrt(r_reg) scale)
            R_in = np.asarray(R_in)
            self.vars["Rx"] = leaf(R_in.real)
            self.vars["Ry"] = leaf(R_in.imag)
            Rx, Ry = self.vars["Rx"], self.vars["Ry"]
        else:                                         # model.py:36-39
            self.vars["Rx"] = leaf(raw["Rx"])
            self.vars["Ry"] = leaf(raw["Ry"])
            s_r = self._rsqrt_const(self.r_reg)
            Rx = s_r * self.vars["Rx"]
            Ry = s_r * self.vars["Ry"]
        R = torch.complex(Rx, Ry)                     # model.py:41
        # model.py:42 -- broadcast quirk: subtracts the diagonal VECTOR along the last axis,
        # R_eff[i,j] = R[i,j] - R[j,j]
        self.R = R - torch.diagonal(R)

        if freqs_in is not None:                      # model.py:44-46
            self.vars["freqs"] = leaf(freqs_in)
            self.freqs = self.vars["freqs"]
        else:                                         # model.py:49-50
            self.vars["freqs"] = leaf(raw["freqs"])
            self.freqs = self._rsqrt_const(self.h_reg) * self.vars["freqs"]

    def _rsqrt_const(self, v: float):
        """tf.rsqrt(python float) evaluates in float32 (model.py:36,49)."""
        r32 = np.float32(1.0) / np.sqrt(np.float32(v))
        return torch.tensor(float(r32), dtype=self.rdt)

    # phases = exp(1j * freqsc * t), t float32 (model.py:304-305, 321-322, 176-178)
    def _phases(self, t32: np.float32):
        if self.mode == "f32":
            ang = self.freqs * torch.tensor(t32, dtype=torch.float32)
        else:
            # keep the reference's definition: the angle is the float32 product.
            # straight-through so d(angle)/d(freqs) = t, as TF differentiates it.
            exact = self.freqs * float(t32)
            rounded = (self.freqs.detach().to(torch.float32) * torch.tensor(t32, dtype=torch.float32)).to(torch.float64)
            ang = exact + (rounded - exact.detach())
        return torch.complex(torch.cos(ang), torch.sin(ang))


# ----------------------------------------------------------------------------
# PsiCMPS (model.py:206-334)
# ----------------------------------------------------------------------------
class PsiCMPSOracle(CMPSOracle):
    def __init__(self, hp: HP, raw: Optional[dict] = None, mode: str = "f32", **kw):
        super().__init__(hp, raw, mode=mode, **kw)
        raw = raw or {}
        D = self.bond_d
        rg = self.vars["A"].requires_grad
        if "psi_x" in raw:
            px, py = raw["psi_x"], raw["psi_y"]
        else:  # deterministic stand-in for the glorot default (model.py:218-219)
            rng = np.random.default_rng(12345)
            lim = math.sqrt(3.0 / D)
            px = rng.uniform(-lim, lim, D)
            py = rng.uniform(-lim, lim, D)
        self.vars["psi_x"] = torch.tensor(np.asarray(px, np.float32), dtype=self.rdt, requires_grad=rg)
        self.vars["psi_y"] = torch.tensor(np.asarray(py, np.float32), dtype=self.rdt, requires_grad=rg)
        psi = torch.complex(self.vars["psi_x"], self.vars["psi_y"])   # model.py:221
        self.psi_0 = self._normalize_psi(psi)                         # model.py:222 (axis=None)

    # model.py:327-334
    def _normalize_psi(self, x, axis=None, epsilon=1e-12):
        sq = torch.square(torch.abs(x))
        square_sum = sq.sum() if axis is None else sq.sum(dim=axis, keepdim=True)
        inv = torch.rsqrt(torch.clamp(square_sum, min=epsilon))
        return x * inv.to(self.cdt)

    # model.py:300-317
    def _update_ancilla_psi(self, psi, signal, t32):
        signal = (signal / self.A).to(self.cdt)                       # :303
        phases = self._phases(t32)                                    # :304-305
        Upsi = psi * torch.conj(phases)                               # :306
        Rdag = self.R.conj().transpose(0, 1)                          # :308
        RUpsi = torch.einsum("bc,ac->ab", self.R, Upsi)               # :309
        RdagRUpsi = torch.einsum("bc,ac->ab", Rdag, RUpsi)            # :310
        delta_Upsi = -self.delta_t * self.sigma ** 2 * RdagRUpsi / 2.0  # :312
        delta_Upsi = delta_Upsi + signal.unsqueeze(1) * RUpsi         # :313
        delta_psi = phases * delta_Upsi                               # :315
        return psi + delta_psi                                        # :317

    # model.py:319-325
    def _expectation(self, psi, t32):
        phases = self._phases(t32)
        Upsi = psi * torch.conj(phases)
        exp = torch.einsum("ab,bc,ac->a", torch.conj(Upsi), self.R, Upsi)
        return 2 * exp.real

    # model.py:293-294
    def _inc_loss_psi(self, psi, signal, t32):
        return -torch.log(1.0 + self._expectation(psi, t32) * signal / self.A)

    def _as_data(self, data):
        return torch.as_tensor(np.asarray(data, dtype=np.float32)).to(self.rdt)

    def _incs(self, data):
        data = self._as_data(data)
        return (data[:, 1:] - data[:, :-1]).transpose(0, 1)           # model.py:263-264

    # model.py:257-267, 276-282 ; returns loss[B] BEFORE the reduce_mean
    def loss_per_clip(self, data):
        incs = self._incs(data)
        B = incs.shape[1]
        psi = self.psi_0.unsqueeze(0).repeat(B, 1)
        loss = torch.zeros(B, dtype=self.rdt)
        t = np.float32(0.0)
        for k in range(incs.shape[0]):
            sig = incs[k]
            psi = self._update_ancilla_psi(psi, sig, t)               # :278
            loss = loss + self._inc_loss_psi(psi, sig, t)             # :279
            psi = self._normalize_psi(psi, axis=1)                    # :280
            t = np.float32(t + self.dt32)                             # :281
        return loss

    def loss_and_abs_terms(self, data):
        """(loss_b, sum_k |term_{k,b}|): the second is the condition of the sum -- a clip whose terms
        cancel (loss_b << sum |terms|) cannot be held to a relative tolerance on loss_b alone."""
        incs = self._incs(data)
        B = incs.shape[1]
        psi = self.psi_0.unsqueeze(0).repeat(B, 1)
        loss = torch.zeros(B, dtype=self.rdt)
        absum = torch.zeros(B, dtype=self.rdt)
        t = np.float32(0.0)
        with torch.no_grad():
            for k in range(incs.shape[0]):
                sig = incs[k]
                psi = self._update_ancilla_psi(psi, sig, t)
                term = self._inc_loss_psi(psi, sig, t)
                loss = loss + term
                absum = absum + term.abs()
                psi = self._normalize_psi(psi, axis=1)
                t = np.float32(t + self.dt32)
        return loss, absum

    def loss(self, data):
        return self.loss_per_clip(data).mean()                        # :267

    # model.py:231-240, 269-274 ; [B, T-1, D]
    def psi_evolve_with_data(self, data):
        incs = self._incs(data)
        B = incs.shape[1]
        psi = self.psi_0.unsqueeze(0).repeat(B, 1)
        out = []
        t = np.float32(0.0)
        for k in range(incs.shape[0]):
            psi = self._update_ancilla_psi(psi, incs[k], t)
            psi = self._normalize_psi(psi, axis=1)
            t = np.float32(t + self.dt32)
            out.append(psi)
        return torch.stack(out, dim=1)

    # model.py:242-251, 284-291 with the noise tensor supplied ([length, n], already scaled
    # by sigma*sqrt(temp*delta_t), model.py:246)
    def sample_from_noise(self, noise):
        noise = torch.as_tensor(np.asarray(noise, dtype=np.float32)).to(self.rdt)
        L, n = noise.shape
        psi = self.psi_0.unsqueeze(0).repeat(n, 1)
        sample = torch.zeros(n, dtype=self.rdt)
        t = np.float32(0.0)
        outs = []
        for k in range(L):
            increment = self._expectation(psi, t) * self.delta_t + noise[k]   # :286
            sample = sample + increment                                        # :287
            psi = self._update_ancilla_psi(psi, increment, t)                  # :288
            psi = self._normalize_psi(psi, axis=1)                             # :289
            t = np.float32(t + self.dt32)                                      # :290
            outs.append(sample)
        return self.A * torch.stack(outs, dim=1)                               # :251  [n, L]

    def sample(self, num_samples, length, temp=1.0, rng=None):
        rng = rng or np.random.default_rng(0)
        noise = (rng.standard_normal((length, num_samples)) *
                 (self.sigma * np.sqrt(temp * self.delta_t))).astype(np.float32)
        return self.sample_from_noise(noise)


# ----------------------------------------------------------------------------
# RhoCMPS (model.py:55-203)
# ----------------------------------------------------------------------------
class RhoCMPSOracle(CMPSOracle):
    def __init__(self, hp: HP, raw: Optional[dict] = None, W_in=None, mode: str = "f32", **kw):
        super().__init__(hp, raw, mode=mode, **kw)
        raw = raw or {}
        D = self.bond_d
        rg = self.vars["A"].requires_grad
        self.rank_rho_0 = hp.initial_rank if hp.initial_rank is not None else D  # :62-65
        if W_in is not None:                                                     # :119-123
            W_in = np.asarray(W_in)
            wx, wy = W_in.real, W_in.imag
        elif "Wx" in raw:
            wx, wy = raw["Wx"], raw["Wy"]
        else:
            rng = np.random.default_rng(54321)
            lim = math.sqrt(6.0 / (self.rank_rho_0 + D))
            wx = rng.uniform(-lim, lim, (self.rank_rho_0, D))
            wy = rng.uniform(-lim, lim, (self.rank_rho_0, D))
        self.vars["Wx"] = torch.tensor(np.asarray(wx, np.float32), dtype=self.rdt, requires_grad=rg)
        self.vars["Wy"] = torch.tensor(np.asarray(wy, np.float32), dtype=self.rdt, requires_grad=rg)
        W = torch.complex(self.vars["Wx"], self.vars["Wy"])                      # :127
        rho_0 = W.conj().transpose(0, 1) @ W                                     # :128
        self.rho_0 = rho_0 / torch.einsum("ii->", rho_0)                         # :129

    def _Rt(self, t32):
        phases = self._phases(t32)                                               # :178
        return torch.einsum("a,ab,b->ab", phases, self.R, torch.conj(phases))    # :179

    # model.py:172-187
    def _update_ancilla_rho(self, rho, signal, t32):
        signal = (signal / self.A).to(self.cdt)
        batch = rho.shape[0]
        Rt = self._Rt(t32)
        RR_dag = (Rt.conj().transpose(0, 1) @ Rt).unsqueeze(0)                   # :180-181
        IR = torch.einsum("a,bc->abc", signal, Rt)                               # :182
        one = torch.eye(self.bond_d, dtype=self.cdt).unsqueeze(0).repeat(batch, 1, 1)
        U = one + (-0.5 * RR_dag * self.delta_t * self.sigma ** 2 + IR)          # :184
        U_dag = U.conj().transpose(1, 2)
        return torch.einsum("abc,acd,ade->abe", U, rho, U_dag)                   # :186

    # model.py:189-196
    def _expectation(self, rho, t32):
        Rt = self._Rt(t32)
        x = Rt + Rt.conj().transpose(0, 1)
        exp = torch.einsum("ab,cba->c", x, rho)   # trace(einsum('ab,cbd->cad'))
        return exp.real

    # model.py:198-203
    def _normalize_rho(self, x, epsilon=1e-12):
        tr = torch.einsum("aii->a", x).reshape(-1, 1, 1)
        inv = torch.reciprocal(torch.clamp(tr.real, min=epsilon))
        return x * inv.to(self.cdt)

    def _inc_loss_rho(self, rho, signal, t32):                                   # :169-170
        return -torch.log(1.0 + self._expectation(rho, t32) * signal / self.A)

    def _as_data(self, data):
        return torch.as_tensor(np.asarray(data, dtype=np.float32)).to(self.rdt)

    def _incs(self, data):
        data = self._as_data(data)
        return (data[:, 1:] - data[:, :-1]).transpose(0, 1)

    # model.py:132-142, 152-158
    def loss_per_clip(self, data):
        incs = self._incs(data)
        B = incs.shape[1]
        rho = self.rho_0.unsqueeze(0).repeat(B, 1, 1)
        loss = torch.zeros(B, dtype=self.rdt)
        t = np.float32(0.0)
        for k in range(incs.shape[0]):
            rho = self._update_ancilla_rho(rho, incs[k], t)
            loss = loss + self._inc_loss_rho(rho, incs[k], t)
            rho = self._normalize_rho(rho)
            t = np.float32(t + self.dt32)
        return loss

    def loss(self, data):
        return self.loss_per_clip(data).mean()

    # model.py:76-85, 144-150 ; [B, T-1, D, D]
    def rho_evolve_with_data(self, data):
        incs = self._incs(data)
        B = incs.shape[1]
        rho = self.rho_0.unsqueeze(0).repeat(B, 1, 1)
        t = np.float32(0.0)
        out = []
        for k in range(incs.shape[0]):
            rho = self._update_ancilla_rho(rho, incs[k], t)
            rho = self._normalize_rho(rho)
            t = np.float32(t + self.dt32)
            out.append(rho)
        return torch.stack(out, dim=1)

    # model.py:160-167 ; returns (rho trajectory [n,L,D,D], samples [n,L] BEFORE the A scale)
    def _sample_scan(self, noise):
        noise = torch.as_tensor(np.asarray(noise, dtype=np.float32)).to(self.rdt)
        L, n = noise.shape
        rho = self.rho_0.unsqueeze(0).repeat(n, 1, 1)
        sample = torch.zeros(n, dtype=self.rdt)
        t = np.float32(0.0)
        rhos, outs = [], []
        for k in range(L):
            increment = self._expectation(rho, t) * self.delta_t + noise[k]
            sample = sample + increment
            rho = self._update_ancilla_rho(rho, increment, t)
            rho = self._normalize_rho(rho)
            t = np.float32(t + self.dt32)
            rhos.append(rho)
            outs.append(sample)
        return torch.stack(rhos, dim=1), torch.stack(outs, dim=1)

    def sample_from_noise(self, noise):                                           # :103-112
        _, s = self._sample_scan(noise)
        return self.A * s

    def rho_evolve_with_sampling_from_noise(self, noise):                         # :87-93
        r, _ = self._sample_scan(noise)
        return r

    def purity_from_noise(self, noise):                                           # :95-101
        r, _ = self._sample_scan(noise)
        return torch.einsum("abcd,abdc->ab", r, r).real

    def make_noise(self, num_samples, length, temp=1.0, rng=None):
        rng = rng or np.random.default_rng(0)
        return (rng.standard_normal((length, num_samples)) *
                (self.sigma * np.sqrt(temp * self.delta_t))).astype(np.float32)


# ----------------------------------------------------------------------------
# train.py:55-60 regularised total loss
# ----------------------------------------------------------------------------
def total_loss(model: CMPSOracle, data):
    h_l2sqnorm = torch.sum(torch.square(model.freqs))
    r_l2sqnorm = torch.sum(torch.conj(model.R) * model.R).real
    return model.loss(data) + model.hp.h_reg * h_l2sqnorm + model.hp.r_reg * r_l2sqnorm


def grads_of(model: CMPSOracle, scalar) -> dict:
    """d scalar / d raw variables, the set tf.train.AdamOptimizer.minimize differentiates (train.py:89)."""
    names = list(model.vars.keys())
    gs = torch.autograd.grad(scalar, [model.vars[n] for n in names], allow_unused=True)
    return {n: (g.detach().numpy() if g is not None else None) for n, g in zip(names, gs)}
