"""The reference's own 12 tests (/root/reference/tests/test_model.py, test_data.py) ported onto the
oracle -- the structural properties are the only pins the reference holds for this path."""
import numpy as np
import torch

from oracle.cmps_oracle import (HP, CMPSOracle, PsiCMPSOracle, RhoCMPSOracle, damped_sine,
                                random_raw_params, ref_test_hparams)

SAMPLE_DURATION = 2 ** 8   # tests/test_model.py:9


def _setup(rho=False, seed=0):
    hp = ref_test_hparams()
    raw = random_raw_params(hp, np.random.default_rng(seed), rho=rho)
    data = damped_sine(hp.minibatch_size, SAMPLE_DURATION, hp.delta_t, np.random.default_rng(seed + 1))
    return hp, raw, data


def test_R_has_no_diagonal_elements():           # tests/test_model.py:19-25
    hp, raw, _ = _setup()
    m = CMPSOracle(hp, raw)
    np.testing.assert_allclose(np.diag(m.R.detach().numpy()), np.zeros(hp.bond_dim), atol=1e-7)
    # and the broadcast quirk (model.py:42): every column is shifted by its diagonal element
    R = (m._rsqrt_const(hp.r_reg).numpy() * (raw["Rx"] + 1j * raw["Ry"])).astype(np.complex64)
    np.testing.assert_allclose(m.R.detach().numpy(), R - np.diag(R)[None, :], rtol=1e-5, atol=1e-4)


def test_rho_loss_not_nan():                     # :33-39
    hp, raw, data = _setup(rho=True)
    assert np.isfinite(float(RhoCMPSOracle(hp, raw).loss(data)))


def test_rho0_is_a_density_matrix():             # :41-48
    hp, raw, _ = _setup(rho=True)
    r = RhoCMPSOracle(hp, raw).rho_0.detach().numpy()
    np.testing.assert_allclose(r, r / np.trace(r), rtol=1e-6)
    np.testing.assert_allclose(r, r.conj().T, rtol=1e-6, atol=1e-7)


def test_rho_evolved_with_data_remains_normalized():   # :50-57
    hp, raw, data = _setup(rho=True)
    out = RhoCMPSOracle(hp, raw).rho_evolve_with_data(data).detach().numpy()
    np.testing.assert_allclose(np.trace(out, axis1=-2, axis2=-1), np.ones(out.shape[:2]), rtol=1e-5)


def test_rho_evolved_sampling_remains_normalized():    # :59-67
    hp, raw, _ = _setup(rho=True)
    m = RhoCMPSOracle(hp, raw)
    out = m.rho_evolve_with_sampling_from_noise(m.make_noise(5, SAMPLE_DURATION)).detach().numpy()
    assert out.shape[:2] == (5, SAMPLE_DURATION)
    np.testing.assert_allclose(np.trace(out, axis1=-2, axis2=-1), np.ones((5, SAMPLE_DURATION)), rtol=1e-4)


def test_rho_trivial_update_of_ancilla():              # :69-83
    hp = ref_test_hparams()
    m = RhoCMPSOracle(hp, None, freqs_in=np.zeros(hp.bond_dim, np.float32),
                      R_in=np.zeros((hp.bond_dim,) * 2, np.complex64))
    signal = torch.tensor(np.random.default_rng(0).random(hp.minibatch_size).astype(np.float32))
    stack = m.rho_0.unsqueeze(0).repeat(hp.minibatch_size, 1, 1)
    np.testing.assert_allclose(m._update_ancilla_rho(stack, signal, np.float32(0.)).detach().numpy(),
                               stack.detach().numpy(), rtol=1e-6)


def _qubit_hp(r_reg):
    return HP(minibatch_size=8, bond_dim=2, delta_t=1 / 16000, sigma=1, initial_rank=None, A=1.,
              h_reg=2 / (np.pi * 16000) ** 2, r_reg=r_reg)


def test_rho_sampling_two_level_system():              # :85-103
    R = np.array([[0, 1], [0, 0]], dtype=np.complex64)
    freqs = np.array([10, -10], dtype=np.float32)
    q = RhoCMPSOracle(_qubit_hp(2 / (np.pi * 16000)), None, R_in=R, freqs_in=freqs)
    w = q.sample_from_noise(q.make_noise(2, 512)).detach().numpy()
    assert w.shape == (2, 512) and np.all(np.isfinite(w))


def test_psi_loss_not_nan():                           # :107-113
    hp, raw, data = _setup()
    assert np.isfinite(float(PsiCMPSOracle(hp, raw).loss(data)))


def test_psi_evolved_with_data_remains_normalized():   # :115-122
    hp, raw, data = _setup()
    out = PsiCMPSOracle(hp, raw).psi_evolve_with_data(data).detach().numpy()
    np.testing.assert_allclose(np.linalg.norm(out, axis=-1), np.ones(out.shape[:2]), rtol=1e-5)


def test_psi_trivial_update_of_ancilla():              # :124-138
    hp = ref_test_hparams()
    m = PsiCMPSOracle(hp, None, freqs_in=np.zeros(hp.bond_dim, np.float32),
                      R_in=np.zeros((hp.bond_dim,) * 2, np.complex64))
    signal = torch.tensor(np.random.default_rng(0).random(hp.minibatch_size).astype(np.float32))
    stack = m.psi_0.unsqueeze(0).repeat(hp.minibatch_size, 1)
    np.testing.assert_allclose(m._update_ancilla_psi(stack, signal, np.float32(0.)).detach().numpy(),
                               stack.detach().numpy(), rtol=1e-6)


def test_psi_sampling_two_level_system():              # :140-158
    R = np.array([[0, 1], [0, 0]], dtype=np.complex64)
    freqs = np.array([10, -10], dtype=np.float32)
    q = PsiCMPSOracle(_qubit_hp(2 / (np.pi * 16000) ** 2), None, R_in=R, freqs_in=freqs)
    w = q.sample(num_samples=2, length=512).detach().numpy()
    assert w.shape == (2, 512) and np.all(np.isfinite(w))


def test_get_audio_correct_shape():                    # tests/test_data.py:12-16
    hp = HP(minibatch_size=8, bond_dim=8, delta_t=0.001)
    assert damped_sine(hp.minibatch_size, SAMPLE_DURATION, hp.delta_t, np.random.default_rng(0)).shape == \
        (hp.minibatch_size, SAMPLE_DURATION)


def test_oracle_gradient_against_finite_differences():
    """finite-difference spot check of the oracle's gradient (float64 mode, central
    differences along random directions of Rx, Ry, psi_x, psi_y, A; freqs is excluded because the
    reference's float32 phase-angle rounding makes the loss piecewise in it).  The raw variables are
    float32 values, so the perturbed points are rounded to float32 and the exact displacement is used."""
    from oracle.cmps_oracle import grads_of, total_loss
    hp = HP(bond_dim=4, minibatch_size=2)
    rng = np.random.default_rng(21)
    raw = random_raw_params(hp, rng)
    data = damped_sine(2, 60, hp.delta_t, np.random.default_rng(22))
    o = PsiCMPSOracle(hp, raw, mode="f64")
    g = grads_of(o, total_loss(o, data))
    for name in ("Rx", "Ry", "psi_x", "psi_y", "A"):
        base = np.asarray(raw.get(name, hp.A), np.float32)
        v = rng.standard_normal(base.shape)
        eps = 1e-3 * (np.abs(base).max() + 1e-3)
        plus = (base.astype(np.float64) + eps * v).astype(np.float32)
        minus = (base.astype(np.float64) - eps * v).astype(np.float32)
        lp = float(total_loss(PsiCMPSOracle(hp, {**raw, name: plus}, mode="f64", requires_grad=False), data))
        lm = float(total_loss(PsiCMPSOracle(hp, {**raw, name: minus}, mode="f64", requires_grad=False), data))
        an = float(np.sum(np.asarray(g[name]) * (plus.astype(np.float64) - minus.astype(np.float64))))
        assert abs((lp - lm) - an) <= 2e-5 * max(abs(an), abs(lp - lm)) + 1e-12, (name, lp - lm, an)
