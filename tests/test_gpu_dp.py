"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): two NCCL ranks, each with half of the batch,
must produce the same parameter update as one GPU on the full batch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, native=False):
    import torch.distributed as dist
    from audio_mps_b200 import HParams, PsiCMPS, damped_sine
    from audio_mps_b200.train import Trainer, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    hp = HParams(minibatch_size=6, bond_dim=8, delta_t=1 / 16000, sigma=0.0001,
                 h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
    data = damped_sine(6, 700, hp.delta_t, np.random.default_rng(1))
    model = PsiCMPS(hp, device=dev, seed=0)
    tr = Trainer(model, native_comm=native)      # native: the C ABI's own NCCL communicator
    lo, hi = shard_bounds(6, rank, world)
    losses = [float(tr.step(data[lo:hi], global_batch=6)) for _ in range(3)]
    q.put((rank, losses, {n: p.detach().cpu().numpy() for n, p in model.named_parameters()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("native", [False, True])
def test_two_gpu_training_matches_single_gpu(lib, native):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from audio_mps_b200 import HParams, PsiCMPS, damped_sine
    from audio_mps_b200.train import Trainer
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, native)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, losses, params = q.get(timeout=300)
        got[r] = (losses, params)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hp = HParams(minibatch_size=6, bond_dim=8, delta_t=1 / 16000, sigma=0.0001,
                 h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100., learning_rate=0.001)
    data = damped_sine(6, 700, hp.delta_t, np.random.default_rng(1))
    model = PsiCMPS(hp, device=torch.device("cuda", 0), seed=0)
    tr = Trainer(model)
    ref_losses = [float(tr.step(data, global_batch=6)) for _ in range(3)]
    np.testing.assert_allclose(got[0][0], got[1][0], rtol=0, atol=0)       # same all-reduced loss on both ranks
    np.testing.assert_allclose(got[0][0], ref_losses, rtol=2e-5)
    for n, p in model.named_parameters():
        np.testing.assert_array_equal(got[0][1][n], got[1][1][n])          # replicas stay identical
        np.testing.assert_allclose(got[0][1][n], p.detach().cpu().numpy(), rtol=1e-4, atol=1e-6)
