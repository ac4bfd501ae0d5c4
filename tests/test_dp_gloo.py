"""world_size-2 gloo test of the data-parallel host logic: clips sharded contiguously,
w_b = 1/B_global, ONE all-reduce(sum) of the packed effective-parameter gradient, identical update
on every rank.  The per-shard gradient comes from the oracle here (CPU box, no GPU) -- the kernel's
packed layout [gR | gf | gpsi0 | gA | sum w_b loss_b] is what is exchanged."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_mps_b200.train import shard_bounds
from oracle.cmps_oracle import HP, PsiCMPSOracle, damped_sine, random_raw_params


def _packed(o, data, w):
    lpc = o.loss_per_clip(data)
    tot = (lpc * torch.tensor(w, dtype=lpc.dtype)).sum()
    gR, gf, gp, gA = torch.autograd.grad(tot, [o.R, o.freqs, o.psi_0, o.A])
    return torch.cat([torch.view_as_real(gR).reshape(-1), gf, torch.view_as_real(gp).reshape(-1),
                      gA.reshape(1), tot.detach().reshape(1)]).to(torch.float32)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hp = HP(bond_dim=4, minibatch_size=6)
    raw = random_raw_params(hp, np.random.default_rng(0))          # replicated parameters
    data = damped_sine(6, 120, hp.delta_t, np.random.default_rng(1))
    lo, hi = shard_bounds(6, rank, world)
    packed = _packed(PsiCMPSOracle(hp, raw, mode="f64"), data[lo:hi], np.full(hi - lo, 1.0 / 6))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    q.put((rank, packed.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_batch():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hp = HP(bond_dim=4, minibatch_size=6)
    raw = random_raw_params(hp, np.random.default_rng(0))
    data = damped_sine(6, 120, hp.delta_t, np.random.default_rng(1))
    full = _packed(PsiCMPSOracle(hp, raw, mode="f64"), data, np.full(6, 1.0 / 6)).numpy()
    np.testing.assert_allclose(got[0], got[1], rtol=0, atol=0)     # every rank holds the same sum
    np.testing.assert_allclose(got[0], full, rtol=2e-5, atol=1e-7)
    assert got[0].shape == (2 * 16 + 3 * 4 + 2,)
