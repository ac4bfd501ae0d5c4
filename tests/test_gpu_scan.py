"""-m gpu: the parallel-in-time tensor-core scan (tcgen05 operator composition, amps_psi_loss_fwd_scan)
must reproduce the sequential kernel's per-clip loss and the oracle's."""
import numpy as np
import pytest
import torch

from audio_mps_b200 import PsiCMPS
from oracle.cmps_oracle import PsiCMPSOracle, damped_sine, random_raw_params
from tests.util import hp_pair, rel, set_raw

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,B,T", [(64, 1, 6000), (64, 3, 2500), (32, 2, 4000), (7, 1, 1200), (64, 1, 40)])
def test_scan_matches_sequential(cuda, lib, D, B, T):
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(ohp, np.random.default_rng(D))
    data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(T))
    m = PsiCMPS(php, device=cuda)
    set_raw(m, raw)
    seq = m.loss_per_clip(data).detach().cpu().numpy()
    scan = m.loss_per_clip_scan(data).cpu().numpy()
    assert np.all(np.isfinite(scan))
    assert rel(scan, seq) <= 1e-4, (scan, seq)


def test_scan_matches_oracle(cuda, lib):
    D, B, T = 64, 2, 700
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(ohp, np.random.default_rng(1))
    data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(2))
    m = PsiCMPS(php, device=cuda)
    set_raw(m, raw)
    ref = PsiCMPSOracle(ohp, raw, mode="f64").loss_per_clip(data).detach().numpy()
    assert rel(m.loss_per_clip_scan(data).cpu().numpy(), ref) <= 1e-4
