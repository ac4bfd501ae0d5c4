"""-m gpu: round-2 ABI / parity cases -- self-contained workspaces (no context-cached tables), the
zero-state clamp of model.py:331-333, the fused parameter chain + regulariser the Trainer uses, and
"no device allocation after amps_create" on the device entry points."""
import ctypes as C

import numpy as np
import pytest
import torch

from audio_mps_b200 import HParams, PsiCMPS, RhoCMPS, _lib
from audio_mps_b200.train import Trainer
from oracle.cmps_oracle import PsiCMPSOracle, RhoCMPSOracle, damped_sine, grads_of, random_raw_params, total_loss
from tests.util import hp_pair, rel, rel_clip, relc, set_raw

pytestmark = pytest.mark.gpu


def test_two_models_with_different_delta_t_interleaved(cuda, lib):
    """fwd(A, dt1) -> fwd(B, dt2) -> bwd(A) -> bwd(B): each backward must see ITS OWN float32 time
    table (it lives in the caller's workspace, not in the context)."""
    D, B, T = 8, 3, 700
    models, refs, losses = [], [], []
    for dt, seed in ((1 / 16000, 0), (1 / 11025, 5)):
        ohp, php = hp_pair(bond_dim=D, minibatch_size=B, delta_t=dt)
        raw = random_raw_params(ohp, np.random.default_rng(seed))
        data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(seed + 1))
        m = PsiCMPS(php, device=cuda)
        set_raw(m, raw)
        o = PsiCMPSOracle(ohp, raw, mode="f64")
        refs.append(grads_of(o, o.loss_per_clip(data).mean()))
        models.append(m)
        losses.append(m.loss_per_clip(data))          # both forwards first
    for m, l, gref in zip(models, losses, refs):      # then both backwards
        l.mean().backward()
        for n in ("Rx", "Ry", "freqs_raw", "psi_x", "psi_y", "A"):
            r = gref["freqs" if n == "freqs_raw" else n]
            assert rel(getattr(m, n).grad.cpu().numpy(), r) <= 1e-3, n


def test_rho_two_models_interleaved(cuda, lib):
    D, B, T = 4, 2, 200
    models, refs, losses = [], [], []
    for dt, seed in ((1 / 16000, 0), (1 / 8000, 3)):
        ohp, php = hp_pair(bond_dim=D, minibatch_size=B, delta_t=dt)
        raw = random_raw_params(ohp, np.random.default_rng(seed), rho=True)
        data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(seed + 1))
        m = RhoCMPS(php, device=cuda)
        set_raw(m, raw)
        o = RhoCMPSOracle(ohp, raw, mode="f64")
        refs.append(grads_of(o, o.loss_per_clip(data).mean()))
        models.append(m)
        losses.append(m.loss_per_clip(data))
    for m, l, gref in zip(models, losses, refs):
        l.mean().backward()
        for n in ("Rx", "Ry", "freqs_raw", "Wx", "Wy", "A"):
            r = gref["freqs" if n == "freqs_raw" else n]
            assert rel(getattr(m, n).grad.cpu().numpy(), r) <= 1e-3, n


@pytest.mark.parametrize("D", [8, 32, 64, 128])
def test_zero_state_is_finite_like_the_reference(cuda, lib, D):
    """psi_0 = 0: the reference's clamp max(sum|.|^2, 1e-12) (model.py:331-333) keeps everything finite
    (loss 0, gradient 0); so must every kernel family."""
    ohp, php = hp_pair(bond_dim=D, minibatch_size=2)
    raw = random_raw_params(ohp, np.random.default_rng(0))
    raw["psi_x"] = np.zeros(D, np.float32)
    raw["psi_y"] = np.zeros(D, np.float32)
    data = damped_sine(2, 150, ohp.delta_t, np.random.default_rng(1))
    o = PsiCMPSOracle(ohp, raw, mode="f64")
    ref = o.loss_per_clip(data)
    assert float(ref.detach().abs().max()) == 0.0
    m = PsiCMPS(php, device=cuda)
    set_raw(m, raw)
    l = m.loss_per_clip(data)
    assert torch.isfinite(l).all() and float(l.detach().abs().max()) == 0.0
    l.mean().backward()
    for n in ("Rx", "Ry", "freqs_raw", "psi_x", "psi_y", "A"):
        g = getattr(m, n).grad
        assert torch.isfinite(g).all(), n
        assert float(g.abs().max()) == 0.0, n
    assert torch.isfinite(m.sample(2, 64)).all()


@pytest.mark.parametrize("D", [8, 32])
def test_fused_parameter_chain_and_regulariser_vs_oracle(cuda, lib, D):
    """Trainer.step trains through loss_per_clip_and_regulariser (amps_psi_params_fwd/_bwd): value and
    raw-variable gradients of mean(loss) + regulariser against the oracle's total_loss (train.py:55-60)."""
    B, T = 4, 600
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(ohp, np.random.default_rng(7))
    data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(8))
    o = PsiCMPSOracle(ohp, raw, mode="f64")
    tot = total_loss(o, data)
    gref = grads_of(o, tot)
    m = PsiCMPS(php, device=cuda)
    set_raw(m, raw)
    lpc, reg = m.loss_per_clip_and_regulariser(data)
    obj = lpc.mean() + reg
    assert rel(float(obj.detach()), float(tot.detach())) <= 1e-5
    oreg = float(tot.detach()) - float(o.loss_per_clip(data).mean().detach())
    assert rel(float(reg.detach()), oreg) <= 1e-5
    obj.backward()
    for n in ("Rx", "Ry", "freqs_raw", "psi_x", "psi_y", "A"):
        r = gref["freqs" if n == "freqs_raw" else n]
        assert rel(getattr(m, n).grad.cpu().numpy(), r) <= 1e-3, n


def test_device_entry_points_allocate_nothing(cuda, lib):
    """the caller owns every buffer; after amps_create the device entry points allocate
    nothing.  100 training steps + samples must leave the driver's free-memory figure unchanged once
    the framework's caching allocator is warm (the library itself never calls cudaMalloc there)."""
    _, php = hp_pair(bond_dim=16, minibatch_size=4)
    m = PsiCMPS(php, device=cuda)
    tr = Trainer(m)
    x = torch.as_tensor(damped_sine(4, 800, php.delta_t, np.random.default_rng(0)), device=cuda)
    noise = torch.randn(256, 3, device=cuda) * 1e-6
    for _ in range(3):
        tr.step(x)
        m.sample_from_noise(noise)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(cuda)
    for _ in range(100):
        tr.step(x)
    m.sample_from_noise(noise)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(cuda)
    assert free1 == free0, (free0, free1)


@pytest.mark.parametrize("D,K,B,T", [(8, None, 3, 500), (32, 64, 3, 500), (64, None, 3, 500), (128, None, 40, 60),
                                     (64, None, 150, 60)])
def test_cuda_graph_trainer_matches_eager(cuda, lib, D, K, B, T):
    """Trainer(cuda_graph=True) captures the whole step once and replays it: parameters after 6 steps are
    bit-identical to the eager trainer's (same kernels, same order), also with the checkpointed backward
    (which forks to the context's second stream inside the capture) and with the partial-wave pipelines
    (D = 128, 40 clips: programmatic-serialisation launches inside the capture; D = 64, 150 clips: three streams)."""
    _, php = hp_pair(bond_dim=D, minibatch_size=B)
    x = torch.as_tensor(damped_sine(B, T, php.delta_t, np.random.default_rng(4)), device=cuda)
    finals = []
    for graph in (False, True):
        m = PsiCMPS(php, device=cuda, seed=0)
        if K is not None:
            m.checkpoint_every = K
        tr = Trainer(m, cuda_graph=graph)
        for _ in range(6):
            loss = tr.step(x)
        torch.cuda.synchronize()
        assert tr.cuda_graph == graph      # the capture did not fall back to eager
        finals.append(([p.detach().clone() for p in m.parameters()], float(loss), tr.global_step))
    (pe, le, se), (pg, lg, sg) = finals
    assert se == sg == 6 and le == lg
    for a, b in zip(pe, pg):
        assert torch.equal(a, b)


@pytest.mark.parametrize("D,B,T", [(64, 150, 40), (40, 297, 35), (128, 40, 30), (100, 77, 25)])
def test_partial_wave_pipelining_matches_oracle(cuda, lib, D, B, T):
    """Batches just beyond one wave of chain CTAs (D = 64 backward: 148 per wave, forward: 296; D = 128: 37 4-CTA
    clusters) take the launch_waves path: chain of the remainder next to the tensor-core pass of the full waves
    (D = 64: on three streams; D = 128: one side stream, the GEMM launched with programmatic serialisation).  Loss and
    gradients against the float64 oracle, and identical to the unsplit run (AMPS_CKPT_SERIAL disables the split
    only at context creation, so the comparison here is against the oracle and against a B <= 148 sub-batch)."""
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(ohp, np.random.default_rng(5))
    data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(6))
    m = PsiCMPS(php, device=cuda)
    set_raw(m, raw)
    lpc = m.loss_per_clip(data)
    w = torch.linspace(0.5, 1.5, B, device=cuda) / B
    (lpc * w).sum().backward()
    o = PsiCMPSOracle(ohp, raw, mode="f64")
    ref = o.loss_per_clip(data)
    gref = grads_of(o, (ref * torch.as_tensor(w.cpu().numpy(), dtype=torch.float64)).sum())
    assert rel_clip(lpc.detach().cpu().numpy(), ref.detach().numpy(), floor=0.05) <= 1e-4
    for n in ("Rx", "Ry", "freqs_raw", "psi_x", "psi_y", "A"):
        assert rel(getattr(m, n).grad.cpu().numpy(), gref["freqs" if n == "freqs_raw" else n]) <= 1e-3, n
    # a sub-batch alone (one wave, no split) gives the same per-clip losses bit for bit
    nsub = 100 if D <= 64 else 30
    lpc_sub = m.loss_per_clip(data[:nsub])
    assert torch.equal(lpc_sub.detach(), lpc.detach()[:nsub])
