"""-m gpu: C-ABI behaviour -- host-buffer entry point, error codes, edge shapes, size-independent
properties at BASELINE's full clip length."""
import ctypes as C

import numpy as np
import pytest
import torch

from audio_mps_b200 import HParams, PsiCMPS, _lib
from oracle.cmps_oracle import PsiCMPSOracle, damped_sine, random_raw_params
from tests.util import hp_pair, rel, rel_clip, relc, set_raw

pytestmark = pytest.mark.gpu


def test_host_entry_point_matches_oracle(cuda, lib):
    D, B, T = 8, 4, 500
    ohp, _ = hp_pair(bond_dim=D, minibatch_size=B)
    raw = random_raw_params(ohp, np.random.default_rng(0))
    data = damped_sine(B, T, ohp.delta_t, np.random.default_rng(1))
    o = PsiCMPSOracle(ohp, raw, mode="f64")
    lpc = o.loss_per_clip(data)
    gR, gf, gp, gA = torch.autograd.grad(lpc.mean(), [o.R, o.freqs, o.psi_0, o.A])
    R = np.ascontiguousarray(o.R.detach().numpy().astype(np.complex64))
    f = np.ascontiguousarray(o.freqs.detach().numpy().astype(np.float32))
    p0 = np.ascontiguousarray(o.psi_0.detach().numpy().astype(np.complex64))
    hp = _lib.AmpsHostParams(D=D, reserved=0, R=R.ctypes.data, freqs=f.ctypes.data, psi0=p0.ctypes.data,
                             A=float(ohp.A), sigma=float(ohp.sigma), delta_t=float(ohp.delta_t))
    x = np.ascontiguousarray(data)
    loss = np.zeros(B, np.float32)
    grad = np.zeros(int(lib.amps_psi_grad_count(D)), np.float32)
    h = _lib.context(0)
    _lib.check(h, lib.amps_psi_loss_grad_host(h, C.byref(hp), x.ctypes.data, B, T, 1.0 / B,
                                              loss.ctypes.data, grad.ctypes.data))
    assert rel_clip(loss, lpc.detach().numpy()) <= 1e-4
    n = 2 * D * D
    assert relc(grad[:n].reshape(D, D, 2) @ np.array([1, 1j]), gR.numpy()) <= 1e-3
    assert rel(grad[n:n + D], gf.numpy()) <= 1e-3
    assert relc(grad[n + D:n + 3 * D].reshape(D, 2) @ np.array([1, 1j]), gp.numpy()) <= 1e-3
    assert rel(grad[n + 3 * D], float(gA)) <= 1e-3
    assert rel(grad[n + 3 * D + 1], float(lpc.detach().mean())) <= 1e-4


def test_error_codes(cuda, lib):
    _, php = hp_pair(bond_dim=129)                                  # loss / gradient cover D <= 128
    m = PsiCMPS(php, device=cuda)
    with pytest.raises(_lib.AmpsError) as e:
        m.loss_per_clip(np.zeros((2, 16), np.float32))
    assert e.value.code == -2                                       # AMPS_E_UNSUPPORTED
    with pytest.raises(_lib.AmpsError) as e:                        # ... and so does the sampler
        m.sample(2, 16)
    assert e.value.code == -2
    _, php = hp_pair(bond_dim=100)
    m = PsiCMPS(php, device=cuda)
    assert m.sample(2, 16).shape == (2, 16)
    with pytest.raises(_lib.AmpsError) as e:                        # the tensor-core scan stops at D = 64
        m.loss_per_clip_scan(np.zeros((2, 16), np.float32))
    assert e.value.code == -2
    h = _lib.context(0)
    _, php = hp_pair(bond_dim=8)
    m = PsiCMPS(php, device=cuda)
    R = torch.view_as_real(m.R.detach()).contiguous()
    f = m.freqs.detach().contiguous()
    p0 = torch.view_as_real(m.psi_0.detach()).contiguous()
    p = _lib.AmpsParams(D=8, reserved=0, R_dev=R.data_ptr(), freqs_dev=f.data_ptr(),
                        psi0_dev=p0.data_ptr(), rho0_dev=None, A=100.0, sigma=1e-4, delta_t=1 / 16000)
    x = torch.zeros(2, 64, device=cuda)
    loss = torch.zeros(2, device=cuda)
    ws = torch.zeros(1024, dtype=torch.uint8, device=cuda)
    rc = lib.amps_psi_loss_fwd(h, C.byref(p), x.data_ptr(), 2, 64, loss.data_ptr(), ws.data_ptr(), 1024, 1, None)
    assert rc == -3                                                 # AMPS_E_WORKSPACE
    assert b"workspace" in lib.amps_last_error(h)
    rc = lib.amps_psi_loss_fwd(h, C.byref(p), None, 2, 64, loss.data_ptr(), ws.data_ptr(), 1024, 1, None)
    assert rc == -1                                                 # AMPS_E_INVALID
    rc = lib.amps_psi_loss_fwd(h, C.byref(p), x.data_ptr(), 2, 0, loss.data_ptr(), ws.data_ptr(), 1024, 1, None)
    assert rc == -1


def test_edge_shapes(cuda, lib):
    ohp, php = hp_pair(bond_dim=8)
    raw = random_raw_params(ohp, np.random.default_rng(0))
    m = PsiCMPS(php, device=cuda)
    set_raw(m, raw)
    # T = 1: no step at all -> loss 0, gradient 0
    l = m.loss_per_clip(np.zeros((3, 1), np.float32))
    assert l.shape == (3,) and float(l.detach().abs().max()) == 0.0
    l.sum().backward()
    assert float(m.Rx.grad.abs().max()) == 0.0
    # chunk-boundary lengths (32-step chunks): 32, 33, 34, 64, 65 steps
    for T in (33, 34, 35, 65, 66):
        data = damped_sine(2, T, ohp.delta_t, np.random.default_rng(T))
        ref = PsiCMPSOracle(ohp, raw, mode="f64").loss_per_clip(data).detach().numpy()
        assert rel_clip(m.loss_per_clip(data).detach().cpu().numpy(), ref) <= 1e-4, T
    # zero signal: inc = 0 -> every term is -log(1) = 0
    assert float(m.loss_per_clip(np.ones((2, 100), np.float32)).abs().max()) == 0.0
    assert m.sample(0, 16).shape == (0, 16)


def test_full_length_properties(cuda, lib):
    """Size-independent properties at BASELINE's clip length (64000 samples), D = 32:
    clip independence / batch-permutation invariance, linearity of the gradient in the clip
    weights, additivity of the packed loss slot, determinism, unit norm of the trajectory."""
    D, B, T = 32, 6, 64000
    ohp, php = hp_pair(bond_dim=D, minibatch_size=B)
    m = PsiCMPS(php, device=cuda, seed=3)
    m.time_parallel = "never"      # properties of the one-chain-per-clip kernels (the scan regroups the sums)
    x = torch.from_numpy(damped_sine(B, T, ohp.delta_t, np.random.default_rng(9))).to(cuda)
    l1 = m.loss_per_clip(x)
    perm = torch.tensor([3, 0, 5, 1, 4, 2], device=cuda)
    l2 = m.loss_per_clip(x[perm])
    assert torch.equal(l1[perm], l2)                                # bit-exact: clips are independent
    assert torch.equal(m.loss_per_clip(x[:2]), l1[:2])
    assert torch.all(torch.isfinite(l1))

    def packed(w):
        m.zero_grad()
        (m.loss_per_clip(x) * w).sum().backward()
        return m._last_packed.clone()
    wa = torch.tensor([0.3, 0.0, 0.1, 0.0, 0.2, 0.4], device=cuda)
    wb = torch.tensor([0.0, 0.5, 0.0, 0.25, 0.1, 0.0], device=cuda)
    ga, gb, gab = packed(wa), packed(wb), packed(wa + wb)
    assert torch.equal(packed(wa), ga)                              # deterministic reduction order
    scale = gab.abs().max()
    assert float((ga + gb - gab).abs().max() / scale) <= 2e-5       # linear in the weights
    assert abs(float(gab[-1]) - float((l1 * (wa + wb)).sum())) <= 1e-5 * abs(float(gab[-1])) + 1e-6
    tr = m.psi_evolve_with_data(x[:2])
    assert tr.shape == (2, T - 1, D)
    n = torch.linalg.vector_norm(tr, dim=-1)
    assert float((n - 1).abs().max()) <= 1e-5                       # tests/test_model.py:115-122


def test_train_cli_trains_and_resumes(cuda, lib, tmp_path):
    """train.py's flow end to end: flags, hparams override, Adam steps, checkpoint + restore, samples."""
    import json
    from audio_mps_b200 import train_cli
    argv = ["--dataset=damped_sine", "--sample_duration=512", "--hparams=bond_dim=8,minibatch_size=4",
            f"--logdir={tmp_path}", "--steps=6", "--num_samples=2", "--save_checkpoint_secs=0"]
    train_cli.main(argv + ["--visualize", "--save_summaries_steps=4"])
    logdir = tmp_path / "damped_sine" / f"8_{1/16000}_4"
    recs = [json.loads(l) for l in open(logdir / "scalars.jsonl")]
    assert [r["step"] for r in recs] == [1, 2, 3, 4, 5, 6]
    # the summaries of train.py:62-85 as TensorBoard events: scalars, data audio, frequency histogram, waveform images
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    ea = EventAccumulator(str(logdir), size_guidance={"audio": 0, "images": 0, "histograms": 0, "scalars": 0})
    ea.Reload()
    tags = ea.Tags()
    assert {"model_loss", "total_loss", "A", "sigma", "h_l2norm", "r_l2norm", "gr_decay_time"} <= set(tags["scalars"])
    assert "data/0" in tags["audio"] and "data/3" in tags["audio"] and "frequencies" in tags["histograms"]
    assert "data_waveform/0" in tags["images"] and "sample_waveform/1" in tags["images"]
    assert [e.step for e in ea.Audio("data/0")] == [1, 5]
    assert all(np.isfinite(r["total_loss"]) for r in recs)
    assert np.load(logdir / "samples.npy").shape == (2, 512)
    train_cli.main(argv[:4] + ["--steps=2", "--num_samples=0"])          # resumes from model.pt
    recs = [json.loads(l) for l in open(logdir / "scalars.jsonl")]
    assert recs[-1]["step"] == 8
    for mdl in ("rho_mps",):
        train_cli.main(["--mps_model=" + mdl, "--sample_duration=128", "--hparams=bond_dim=4,minibatch_size=2",
                        f"--logdir={tmp_path}/rho", "--steps=3", "--num_samples=0"])
    # TensorFlow-format checkpoint out, and a fresh run initialised from it (by directory)
    train_cli.main(argv[:4] + ["--steps=1", "--num_samples=0", "--save_tf_checkpoint"])
    from audio_mps_b200 import tf_checkpoint as tfc
    prefix = tfc.latest_checkpoint(str(logdir))
    got = tfc.read_tf_checkpoint(prefix)
    assert int(got["global_step"]) == 9 and got["model/Rx"].shape == (8, 8)
    train_cli.main(argv[:3] + [f"--logdir={tmp_path}/fromtf", f"--init_tf_checkpoint={logdir}", "--steps=1",
                               "--num_samples=0"])
    recs2 = [json.loads(l) for l in open(tmp_path / "fromtf" / "damped_sine" / f"8_{1/16000}_4" / "scalars.jsonl")]
    assert recs2[0]["step"] == 10


def test_device_batch_prefetcher(cuda, lib):
    """Double-buffered host -> device staging: batches come out in submission order, intact, and a
    third submit without a next() is refused."""
    from audio_mps_b200 import DeviceBatchPrefetcher
    pf = DeviceBatchPrefetcher(cuda, (3, 1000))
    hosts = [torch.full((3, 1000), float(k)).pin_memory() + torch.arange(1000)[None, :] for k in range(5)]
    pf.submit(hosts[0])
    for k in range(5):
        x = pf.next()
        if k + 1 < 5:
            pf.submit(hosts[k + 1])
        y = (x * 2).sum()                      # consume on the compute stream
        pf.release()
        assert torch.equal(x.cpu(), hosts[k]) and float(y) == float(hosts[k].sum() * 2)
    pf.submit(hosts[0])
    pf.submit(hosts[1])
    with pytest.raises(RuntimeError):
        pf.submit(hosts[2])
    with pytest.raises(RuntimeError):
        DeviceBatchPrefetcher(cuda, (1, 4)).next()
