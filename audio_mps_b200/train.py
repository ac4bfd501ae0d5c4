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

    def __init__(self, model, learning_rate: Optional[float] = None, group=None, native_comm: bool = False,
                 cuda_graph: bool = False):
        self.model = model
        lr = learning_rate if learning_rate is not None else getattr(model.hparams, "learning_rate", 1e-3)
        on_gpu = all(p.is_cuda for p in model.parameters())
        # cuda_graph: the whole step (parameter chain, scan kernels, all-reduce, Adam: ~50 launches) is captured
        # once per batch shape and replayed -- the step no longer depends on the host keeping ahead of the GPU
        self.cuda_graph = bool(cuda_graph) and on_gpu
        self._graphs = {}
        self.graph_launches_per_step = 0
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8,
                                    fused=True if on_gpu else None,      # one launch per step on the GPU
                                    capturable=self.cuda_graph)
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
        if self.world > 1:
            model.set_data_parallel(group if group is not None else dist.group.WORLD, native=native_comm)
        self.global_step = 0
        self.last_reg = None           # regulariser value of the parameters the last step was taken FROM

    def step(self, x_local, global_batch: Optional[int] = None, regularise: bool = True):
        if self.cuda_graph and isinstance(x_local, torch.Tensor) and x_local.is_cuda:
            try:
                return self._step_graphed(x_local, global_batch, regularise)
            except RuntimeError as e:          # capture refused (driver / allocator state): train eagerly instead
                if self._graphs:
                    raise
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({e}); continuing eagerly")
                self.cuda_graph = False
                torch.cuda.synchronize(x_local.device)
        return self._step_eager(x_local, global_batch, regularise)

    def _step_graphed(self, x_local, global_batch, regularise):
        key = (tuple(x_local.shape), x_local.dtype, global_batch, regularise)
        g = self._graphs.get(key)
        if g is None:
            static_x = x_local.detach().clone()
            # the warm-up steps (allocator, lazy initialisation) and the capture pass must not train: parameters
            # and Adam state are restored IN PLACE afterwards (the graph holds pointers to these tensors)
            params = [p for p in self.model.parameters()]
            saved_p = [p.detach().clone() for p in params]
            saved_s = {i: {k: v.clone() for k, v in self.opt.state[p].items() if torch.is_tensor(v)}
                       for i, p in enumerate(params) if self.opt.state.get(p)}
            step0 = self.global_step
            side = torch.cuda.Stream(device=x_local.device)
            side.wait_stream(torch.cuda.current_stream(x_local.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._step_eager(static_x, global_batch, regularise)
            torch.cuda.current_stream(x_local.device).wait_stream(side)
            from . import _lib
            dev_index = x_local.device.index if x_local.device.index is not None else torch.cuda.current_device()
            l0 = _lib.launch_count(dev_index)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._step_eager(static_x, global_batch, regularise)
            # kernels of libaudiomps.so one replay launches (the library counts launches on the host side)
            self.graph_launches_per_step = _lib.launch_count(dev_index) - l0
            with torch.no_grad():
                for i, p in enumerate(params):
                    p.copy_(saved_p[i])
                    for k, v in self.opt.state[p].items():
                        if torch.is_tensor(v):
                            if i in saved_s and k in saved_s[i]:
                                v.copy_(saved_s[i][k])
                            else:
                                v.zero_()
            self.global_step = step0
            g = self._graphs[key] = (graph, static_x, out)
        graph, static_x, out = g
        static_x.copy_(x_local, non_blocking=True)
        graph.replay()
        self.global_step += 1
        return out

    def _step_eager(self, x_local, global_batch: Optional[int] = None, regularise: bool = True):
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
