"""-m gpu: the CUDA path (through the C ABI) against the committed golden vectors, including
BASELINE config[0] (D=8, 8 clips of 1 s at 16 kHz).  Tolerances: BASELINE.json north_star."""
import numpy as np
import pytest
import torch

from audio_mps_b200 import HParams, PsiCMPS, RhoCMPS
from audio_mps_b200.train import regulariser
from tests.golden_util import PSI_CASES, load, psi_case
from tests.util import rel, rel_clip, relc, set_raw

pytestmark = pytest.mark.gpu


def _model(hp, raw, cuda, cls=PsiCMPS, **kw):
    php = HParams(minibatch_size=hp.minibatch_size, bond_dim=hp.bond_dim, delta_t=hp.delta_t,
                  sigma=hp.sigma, h_reg=hp.h_reg, r_reg=hp.r_reg, initial_rank=None, A=hp.A,
                  learning_rate=0.001)
    m = cls(php, device=cuda, **kw)
    set_raw(m, raw)
    return m


@pytest.mark.parametrize("name", PSI_CASES)
def test_psi_loss_and_grads_vs_golden(cuda, lib, name):
    hp, raw, data, g = psi_case(name)
    m = _model(hp, raw, cuda)
    lpc = m.loss_per_clip(data)
    # per-clip log-likelihood: 1e-4 relative against the exact value of the reference's function
    assert rel_clip(lpc.detach().cpu().numpy(), g["loss_f64"]) <= 1e-4
    # ... and no further from the reference's float32 run than that run's own rounding noise allows
    assert rel_clip(lpc.detach().cpu().numpy(), g["loss_f32"]) <= 1e-4 + 1.5 * rel_clip(g["loss_f32"], g["loss_f64"])
    obj = lpc.mean() + regulariser(m)
    names = ["A", "Rx", "Ry", "freqs_raw", "psi_x", "psi_y"]
    gs = torch.autograd.grad(obj, [getattr(m, n) for n in names])
    for n, gg in zip(names, gs):
        ref = g["grad_" + ("freqs" if n == "freqs_raw" else n) + "_f64"]
        assert rel(gg.cpu().numpy(), ref) <= 1e-3, n


def test_qubit_and_sampler_golden(cuda, lib):
    g = load("qubit_sampling")
    from oracle.cmps_oracle import HP
    hp = HP(minibatch_size=8, bond_dim=2, delta_t=1 / 16000, sigma=1, initial_rank=None, A=1.,
            h_reg=2 / (np.pi * 16000) ** 2, r_reg=2 / (np.pi * 16000) ** 2)
    R = np.array([[0, 1], [0, 0]], dtype=np.complex64)
    fr = np.array([10, -10], dtype=np.float32)
    m = _model(hp, {"psi_x": g["psi_x"], "psi_y": g["psi_y"]}, cuda, R_in=R, freqs_in=fr)
    out = m.sample_from_noise(g["noise"]).cpu().numpy()
    assert out.shape == (2, 512)                                  # tests/test_model.py:158
    assert rel(out, g["psi_sample_f64"]) <= 1e-3
    r = _model(hp, {}, cuda, cls=RhoCMPS, W_in=g["W"], R_in=R, freqs_in=fr)
    assert rel(r.sample_from_noise(g["noise"]).cpu().numpy(), g["rho_sample_f64"]) <= 1e-3
    assert rel(r.purity(2, 512, noise=g["noise"]).cpu().numpy(), g["rho_purity_f64"]) <= 1e-3
    for D in (7, 32):
        s = load(f"psi_sample_d{D}")
        from oracle.cmps_oracle import HP as HP2
        hp2 = HP2(bond_dim=D)
        raw = {k[4:]: s[k] for k in s if k.startswith("raw_")}
        m2 = _model(hp2, raw, cuda)
        assert rel(m2.sample_from_noise(s["noise"]).cpu().numpy(), s["sample_f64"]) <= 1e-3


def test_rho_golden(cuda, lib):
    from oracle.cmps_oracle import ref_test_hparams
    g = load("rho_testhp_d7")
    raw = {k[4:]: g[k] for k in g if k.startswith("raw_")}
    m = _model(ref_test_hparams(), raw, cuda, cls=RhoCMPS)
    assert rel_clip(m.loss_per_clip(g["data"]).detach().cpu().numpy(), g["loss_f64"]) <= 1e-4
    tr = m.rho_evolve_with_data(g["data"]).cpu().numpy()
    assert relc(tr[:, -1], g["traj_last"]) <= 1e-4


FULL = [("psi_c1_full", False), ("psi_c4_d64_full_length", False), ("psi_c4_d64_full_length", True),
        ("psi_c3_d128_full_length", False)]


@pytest.mark.parametrize("name,scan", FULL)
def test_psi_full_length_golden(cuda, lib, name, scan):
    """Full-length (64000-sample) clips against the float64 run of the independent C restatement
    (oracle/mint_golden_c1.py): BASELINE config[1] at FULL size (D=32, 64 clips), and 4 clips at the
    bond dimensions of config[4] (64: chain kernels and the tensor-core scan) and config[3] (128:
    row-split cluster kernels).  Per-clip loss 1e-4, effective-parameter gradients 1e-3."""
    from audio_mps_b200.model import _PsiLossFn
    from oracle.cmps_oracle import HP, damped_sine, random_raw_params
    g = load(name)
    D, B, T, seed = int(g["D"]), int(g["B"]), int(g["T"]), int(g["seed"])
    hp = HP(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(hp, np.random.default_rng(seed))
    data = damped_sine(B, T, hp.delta_t, np.random.default_rng(seed + 1))
    assert abs(np.abs(data.astype(np.float64)).sum() - float(g["data_checksum"])) <= 1e-9 * float(g["data_checksum"])
    m = _model(hp, raw, cuda)
    assert relc(m.R.detach().cpu().numpy(), g["R_eff"]) <= 1e-6
    R = torch.view_as_real(m.R.detach()).clone().requires_grad_()
    f = m.freqs.detach().clone().requires_grad_()
    p0 = torch.view_as_real(m.psi_0.detach()).clone().requires_grad_()
    A = m.A.detach().clone().requires_grad_()
    x = torch.as_tensor(data, device=cuda)
    lpc = _PsiLossFn.apply(R, f, p0, A, x, m, scan)
    assert rel_clip(lpc.detach().cpu().numpy(), g["loss_f64"]) <= 1e-4
    gR, gf, gp, gA = torch.autograd.grad(lpc.mean(), [R, f, p0, A])
    assert relc(torch.view_as_complex(gR).cpu().numpy(), g["geff_R"]) <= 1e-3
    assert rel(gf.cpu().numpy(), g["geff_f"]) <= 1e-3
    assert relc(torch.view_as_complex(gp).cpu().numpy(), g["geff_psi0"]) <= 1e-3
    assert rel(gA.cpu().numpy(), g["geff_A"]) <= 1e-3


def _eff_grads(m, x, scan=False, **kw):
    """(per-clip loss, effective-parameter gradients of mean loss) through the autograd bridge."""
    from audio_mps_b200.model import _PsiLossFn
    R = torch.view_as_real(m.R.detach()).clone().requires_grad_()
    f = m.freqs.detach().clone().requires_grad_()
    p0 = torch.view_as_real(m.psi_0.detach()).clone().requires_grad_()
    A = m.A.detach().clone().requires_grad_()
    lpc = _PsiLossFn.apply(R, f, p0, A, x, m, scan, *kw.get("extra", ()))
    gR, gf, gp, gA = torch.autograd.grad(lpc.mean(), [R, f, p0, A])
    return lpc, torch.view_as_complex(gR), gf, torch.view_as_complex(gp), gA


@pytest.mark.parametrize("name", ["psi_c3_batch_t2000", "psi_c4_batch_t4000"])
def test_psi_full_batch_golden(cuda, lib, name):
    """BASELINE config[3] / config[4] at their FULL per-GPU batch (D=128 x 128 clips on the row-split
    cluster kernels; D=64 x 256 clips = more clips than SMs, the multi-wave single-CTA dispatch) on a
    shortened clip, against the float64 C restatement (oracle/mint_golden_r2.py)."""
    from oracle.cmps_oracle import HP, damped_sine, random_raw_params
    g = load(name)
    D, B, T, seed, off = int(g["D"]), int(g["B"]), int(g["T"]), int(g["seed"]), int(g["offset"])
    hp = HP(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(hp, np.random.default_rng(seed))
    data = np.ascontiguousarray(damped_sine(B, 64000, hp.delta_t, np.random.default_rng(seed + 1))[:, off:off + T])
    assert abs(np.abs(data.astype(np.float64)).sum() - float(g["data_checksum"])) <= 1e-9 * float(g["data_checksum"])
    m = _model(hp, raw, cuda)
    lpc, gR, gf, gp, gA = _eff_grads(m, torch.as_tensor(data, device=cuda))
    assert rel_clip(lpc.detach().cpu().numpy(), g["loss_f64"]) <= 1e-4
    assert relc(gR.cpu().numpy(), g["geff_R"]) <= 1e-3
    assert rel(gf.cpu().numpy(), g["geff_f"]) <= 1e-3
    assert relc(gp.cpu().numpy(), g["geff_psi0"]) <= 1e-3
    assert rel(gA.cpu().numpy(), g["geff_A"]) <= 1e-3


def test_psi_c2_sampler_full_size_golden(cuda, lib):
    """BASELINE config[2] at FULL size -- D=32, 256 samples x 64000 steps from a fixed noise tensor --
    against the float64 C restatement (model.py:242-251, 284-291).  The sampler feeds E(psi) back into
    the state for 64000 steps; every stored column of the cumulative output (every 125th), the last
    column and the row sums are held to 1e-3 of each path's own amplitude."""
    from oracle.cmps_oracle import HP, random_raw_params
    from oracle.mint_golden_r2 import sample_noise
    g = load("psi_c2_sample_full")
    D, n, L, seed, stride = int(g["D"]), int(g["n"]), int(g["L"]), int(g["seed"]), int(g["stride"])
    hp = HP(bond_dim=D, minibatch_size=n)
    raw = random_raw_params(hp, np.random.default_rng(seed))
    noise = sample_noise(hp, L, n, seed + 2)
    assert abs(np.abs(noise.astype(np.float64)).sum() - float(g["noise_checksum"])) <= 1e-9 * float(g["noise_checksum"])
    m = _model(hp, raw, cuda)
    out = m.sample_from_noise(noise).cpu().numpy().astype(np.float64)
    assert out.shape == (n, L)
    amp = g["absmax"][:, None]
    err = np.abs(out[:, stride - 1::stride] - g["sub"]) / amp
    assert err.max() <= 1e-3, (np.unravel_index(err.argmax(), err.shape), err.max())
    assert (np.abs(out[:, -1] - g["last"]) / g["absmax"]).max() <= 1e-3
    assert (np.abs(out.sum(axis=1) - g["rowsum"]) / (L * g["absmax"])).max() <= 1e-3
