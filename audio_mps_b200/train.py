"""Training step around the scan (the caller side of the path, /root/reference/train.py:55-94):
regularised total loss, Adam(lr), and batch data-parallelism -- clips sharded across ranks,
ONE all-reduce of the packed kernel gradient (2 D^2 + 3 D + 2 floats) per step."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist


def regulariser(model) -> torch.Tensor:
    """h_reg*|freqs|^2 + r_reg*|R|^2 on the EFFECTIVE parameters (train.py:55-60)."""
    h_l2sqnorm = torch.sum(torch.square(model.freqs))
    R = model.R
    r_l2sqnorm = torch.sum(torch.conj(R) * R).real
    return model.h_reg * h_l2sqnorm + model.r_reg * r_l2sqnorm


def total_loss(model, data=None) -> torch.Tensor:
    return model.loss_fn(data) + regulariser(model)


def shard_bounds(global_batch: int, rank: int, world: int):
    """Contiguous batch shards; the first (global_batch % world) ranks take one extra clip."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Trainer:
    """Adam on the raw variables; every rank applies the identical update (parameters replicated).

    ``step(x_local, global_batch)`` computes sum_{b in shard} loss_b / global_batch; the kernel
    backward all-reduces its packed gradient over ``group`` (model.set_data_parallel), the
    parameter-only regulariser gradient is computed redundantly on every rank.
    """

    def __init__(self, model, learning_rate: Optional[float] = None, group=None, native_comm: bool = False):
        self.model = model
        lr = learning_rate if learning_rate is not None else getattr(model.hparams, "learning_rate", 1e-3)
        on_gpu = all(p.is_cuda for p in model.parameters())
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8,
                                    fused=True if on_gpu else None)      # one launch per step on the GPU
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
        if self.world > 1:
            model.set_data_parallel(group if group is not None else dist.group.WORLD, native=native_comm)
        self.global_step = 0
        self.last_reg = None           # regulariser value of the parameters the last step was taken FROM

    def step(self, x_local, global_batch: Optional[int] = None, regularise: bool = True):
        m = self.model
        if hasattr(m, "loss_per_clip_and_regulariser"):      # Psi: parameter chain + regulariser fused
            lpc, reg = m.loss_per_clip_and_regulariser(x_local)
        else:
            lpc, reg = m.loss_per_clip(x_local), (regulariser(m) if regularise else None)
        gb = global_batch if global_batch is not None else lpc.shape[0] * self.world
        data_term = lpc.sum() / gb
        self.opt.zero_grad(set_to_none=True)
        obj = data_term + reg if regularise else data_term
        self.last_reg = reg.detach() if reg is not None else None
        obj.backward()
        self.opt.step()
        self.global_step += 1
        # slot [-1] of the packed buffer is sum_b w_b loss_b (all-reduced with the gradient)
        model_loss = m._last_packed[-1] if m._last_packed is not None else data_term.detach()
        return model_loss

    def state_dict(self):
        return {"model": self.model.state_dict(), "opt": self.opt.state_dict(),
                "global_step": self.global_step}

    def load_state_dict(self, sd):
        self.model.load_state_dict(sd["model"])
        self.opt.load_state_dict(sd["opt"])
        self.global_step = sd["global_step"]
