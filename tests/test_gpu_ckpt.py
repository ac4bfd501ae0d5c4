"""-m gpu: checkpoint interval K of the training path (amps_psi_loss_fwd_k / amps_psi_loss_bwd_k).

BASELINE north_star: "a hand-written adjoint backward that recomputes from state checkpoints every K
steps" (the reference's own memory wall is the O(T) activation stack of tf.foldl, model.py:265-266).
K = 1 keeps the whole trajectory; K > 1 keeps one state per K steps and replays window by window.  Both
must give the oracle's loss and gradients, in every kernel family (2-CTA cluster chain D <= 32, single-CTA
chain/filler when the batch exceeds the SM pairs, unified D = 64, row-split 4-CTA cluster D = 128)."""
import ctypes as C

import numpy as np
import pytest
import torch

from audio_mps_b200 import PsiCMPS, _lib
from oracle.cmps_oracle import PsiCMPSOracle, damped_sine, grads_of, random_raw_params
from tests.util import hp_pair, rel, rel_clip, set_raw

pytestmark = pytest.mark.gpu

NAMES = ("Rx", "Ry", "freqs_raw", "psi_x", "psi_y", "A")


def _grads(m, data, K):
    m.checkpoint_every = K
    for p in m.parameters():
        p.grad = None
    lpc = m.loss_per_clip(data)
    lpc.mean().backward()
    return lpc.detach().cpu().numpy(), {n: getattr(m, n).grad.detach().cpu().numpy().copy() for n in NAMES}


@pytest.mark.parametrize("D,B,T", [(8, 3, 1500), (32, 4, 1300), (32, 80, 330), (16, 40, 420), (64, 3, 700), (128, 2, 200)])
@pytest.mark.parametrize("K", [32, 256, 1000])
def test_checkpointed_gradient_matches_oracle_and_full_trajectory(cuda, lib, D, B, T, K):
    """T - 1 is not a multiple of K or of the rescale chunk (ragged last window, ragged last chunk).  (16, 40, .):
    37 < B <= 74 -- the forward runs on the 2-CTA cluster kernels, replay and adjoint as a pair on the single-CTA
    family so that both fit the SMs at once (family_of(paired))."""
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(ohp, np.random.default_rng(D + 1))
    data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(2))
    m = PsiCMPS(php, device=cuda)
    set_raw(m, raw)
    l1, g1 = _grads(m, data, 1)
    lk, gk = _grads(m, data, K)
    # same chain kernel from the same state: the forward value is bit-identical for D <= 32; above that the
    # K = 1 forward takes E_k from the tensor-core expectation pass, the checkpoint-only forward from its
    # in-kernel mat-vec (same value to float32 rounding).  The gradient differs only by the order of the
    # float32 sums (per-window partial tiles, summed window by window)
    if D > 32:
        assert rel_clip(lk, l1) <= 1e-5
    else:
        assert np.array_equal(l1, lk)
    for n in NAMES:
        assert rel(gk[n], g1[n]) <= 1e-4, (n, rel(gk[n], g1[n]))
    if B <= 8:
        o = PsiCMPSOracle(ohp, raw, mode="f64")
        ref = o.loss_per_clip(data)
        gref = grads_of(o, ref.mean())
        assert rel_clip(lk, ref.detach().numpy()) <= 1e-4
        for n in NAMES:
            assert rel(gk[n], gref["freqs" if n == "freqs_raw" else n]) <= 1e-3, n


def test_checkpoint_workspace_and_interval(lib):
    """Workspace at BASELINE C4's per-GPU shard (D = 64, 256 clips x 64 000): 17 GB at K = 1, < 1 GB at
    K = 256; global batch 2048 on ONE GPU needs < 10 GB at K = 2048.  K is rounded up to whole rescale chunks."""
    full = lib.amps_psi_workspace_bytes_k(64, 256, 64000, 1)
    assert full == lib.amps_psi_workspace_bytes(64, 256, 64000, 1) and full > 16e9
    assert lib.amps_psi_workspace_bytes_k(64, 256, 64000, 256) < 1e9
    assert lib.amps_psi_workspace_bytes_k(64, 2048, 64000, 2048) < 1e10
    assert lib.amps_psi_workspace_bytes_k(32, 64, 64000, 2048) < 1.7e8     # C1: 2.1 GB at K = 1
    assert lib.amps_psi_ckpt_interval(32, 100) == 128 and lib.amps_psi_ckpt_interval(128, 100) == 112
    assert lib.amps_psi_ckpt_interval(32, 1) == 1 and lib.amps_psi_ckpt_interval(200, 64) == 0
    assert lib.amps_psi_workspace_bytes_k(32, 4, 1000, 0) == 0


def test_checkpointed_abi_direct_and_errors(cuda, lib):
    """amps_psi_loss_fwd_k / _bwd_k through ctypes: packed gradient equal to the K = 1 entry points;
    undersized workspace and K < 1 are refused."""
    D, B, T, K = 16, 5, 900, 64
    h = _lib.context(0)
    g = torch.Generator().manual_seed(0)
    R = (torch.randn(D, D, 2, generator=g) * 2).to(cuda)
    f = (torch.randn(D, generator=g) * 3000).to(cuda)
    p0 = torch.randn(D, 2, generator=g)
    p0 = (p0 / p0.norm()).to(cuda)
    x = torch.as_tensor(damped_sine(B, T, 1 / 16000, np.random.default_rng(3)), device=cuda)
    w = torch.full((B,), 1.0 / B, device=cuda)
    p = _lib.AmpsParams(D=D, reserved=0, R_dev=R.data_ptr(), freqs_dev=f.data_ptr(), psi0_dev=p0.data_ptr(),
                        rho0_dev=None, A=100.0, sigma=1e-4, delta_t=1 / 16000, A_dev=None)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = []
    for k in (1, K):
        nb = lib.amps_psi_workspace_bytes_k(D, B, T, k)
        ws = torch.empty(nb, dtype=torch.uint8, device=cuda)
        loss = torch.empty(B, device=cuda)
        grad = torch.empty(lib.amps_psi_grad_count(D), device=cuda)
        _lib.check(h, lib.amps_psi_loss_fwd_k(h, C.byref(p), x.data_ptr(), B, T, k, loss.data_ptr(), ws.data_ptr(), nb, st))
        _lib.check(h, lib.amps_psi_loss_bwd_k(h, C.byref(p), x.data_ptr(), B, T, k, w.data_ptr(), ws.data_ptr(), nb,
                                              grad.data_ptr(), st))
        torch.cuda.synchronize()
        outs.append((loss.cpu().numpy(), grad.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert rel(outs[1][1][:2 * D * D].reshape(D, 2 * D), outs[0][1][:2 * D * D].reshape(D, 2 * D)) <= 1e-4
    assert rel(outs[1][1][2 * D * D:], outs[0][1][2 * D * D:]) <= 1e-4
    nb = lib.amps_psi_workspace_bytes_k(D, B, T, K)
    ws = torch.empty(nb, dtype=torch.uint8, device=cuda)
    loss = torch.empty(B, device=cuda)
    assert lib.amps_psi_loss_fwd_k(h, C.byref(p), x.data_ptr(), B, T, K, loss.data_ptr(), ws.data_ptr(), nb - 1, st) == -3
    assert lib.amps_psi_loss_fwd_k(h, C.byref(p), x.data_ptr(), B, T, 0, loss.data_ptr(), ws.data_ptr(), nb, st) == -1


def test_checkpoint_auto_policy(cuda, lib):
    """'auto' keeps the trajectory while it fits and checkpoints beyond that."""
    _, php = hp_pair(bond_dim=8, minibatch_size=2)
    m = PsiCMPS(php, device=cuda)
    assert m._checkpoint_interval(8, 2, 1000) == 1
    m.checkpoint_auto_fraction = 1e-9
    assert m._checkpoint_interval(8, 2, 1000) == 2048
    m.checkpoint_every = 77
    assert m._checkpoint_interval(8, 2, 1000) == 77
