"""Drop-in model-level API of the reference's ``model.py`` (CMPS / PsiCMPS / RhoCMPS) on top of the
B200 scan kernels.

Same constructor arguments, attribute names and method names as /root/reference/model.py; PyTorch
tensors replace TF tensors and the time scan (the reference's ``tf.foldl`` / ``tf.scan`` bodies,
model.py:83-84,91-92,108-109,140-141,238-239,247-248,265-266) is ONE call into ``libaudiomps.so``.
The O(D^2) raw->effective parameter chain (model.py:31-52, 127-130, 221-222) stays here in torch
so autograd carries the kernel's effective-parameter gradients back to the raw variables
``Rx, Ry, freqs, psi_x, psi_y | Wx, Wy, A`` (the set ``AdamOptimizer.minimize`` trains,
train.py:89).  There is no CPU path: scan methods raise unless the model lives on a CUDA device
and the extension is built.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib


def _rsqrt32(v: float) -> float:
    """``tf.rsqrt`` of a python float evaluates in float32 (model.py:36,38,49)."""
    return float(np.float32(1.0) / np.sqrt(np.float32(v)))


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _as_device_f32(obj, device) -> torch.Tensor:
    """Accept torch tensors, numpy arrays and any DLPack exporter."""
    if isinstance(obj, torch.Tensor):
        t = obj
    elif hasattr(obj, "__dlpack__"):
        t = torch.from_dlpack(obj)
    else:
        t = torch.as_tensor(np.asarray(obj))
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def _allreduce_packed(model, packed, h, dev):
    """Data parallel: ONE all-reduce (sum) of the packed effective-parameter gradient, either on the
    C ABI's own NCCL communicator (amps_comm_init) or on a torch.distributed group.  Used by the Psi
    AND the Rho backward, so that replicas cannot diverge between the two models."""
    if getattr(model, "_dp_native", False):
        _lib.check(h, _lib.load().amps_allreduce_grads(h, _ptr(packed), packed.numel(), _stream(dev)))
    elif getattr(model, "_dp_group", None) is not None:
        torch.distributed.all_reduce(packed, op=torch.distributed.ReduceOp.SUM, group=model._dp_group)


# --------------------------------------------------------------------------------------------
# autograd bridge: per-clip loss through amps_psi_loss_fwd / amps_psi_loss_bwd
# --------------------------------------------------------------------------------------------
class _PsiLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, R_ri, freqs, psi0_ri, A, x, model, scan=False):
        dev = x.device
        if dev.type != "cuda":
            raise RuntimeError("PsiCMPS scan requires a CUDA device (no CPU fallback)")
        lib = _lib.load()
        h = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        B, T = x.shape
        D = freqs.shape[0]
        need_grad = any(ctx.needs_input_grad[:4])
        R_ri = R_ri.detach().contiguous().float()
        freqs = freqs.detach().contiguous().float()
        psi0_ri = psi0_ri.detach().contiguous().float()
        # A goes down by device pointer: no host read-back (= stream sync) per training step
        # (a private copy: an in-place parameter update between forward and backward must not change
        # the A the backward sees)
        a_dev = A.detach().reshape(1).float().clone()
        p = _lib.AmpsParams(D=D, reserved=0, R_dev=R_ri.data_ptr(), freqs_dev=freqs.data_ptr(),
                            psi0_dev=psi0_ri.data_ptr(), rho0_dev=None, A=0.0,
                            sigma=float(model.sigma), delta_t=float(model.delta_t),
                            A_dev=a_dev.data_ptr())
        # scan=True: the parallel-in-time tensor-core path (small batches, D <= 64)
        # K > 1 (sequential kernels only): keep one state per K steps and recompute in the backward
        K = 1 if (scan or not need_grad) else model._checkpoint_interval(D, B, T)
        if K > 1:
            nbytes = lib.amps_psi_workspace_bytes_k(D, B, T, K)
        else:
            ws_bytes = lib.amps_psi_scan_workspace_bytes if scan else lib.amps_psi_workspace_bytes
            nbytes = ws_bytes(D, B, T, 1 if need_grad else 0)
        if nbytes == 0:
            raise _lib.AmpsError(-2, f"bond dimension {D} is not supported by the Psi "
                                     f"{'tensor-core scan' if scan else 'kernels'}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        if K > 1:
            rc = lib.amps_psi_loss_fwd_k(h, C.byref(p), _ptr(x), B, T, K, _ptr(loss), _ptr(ws), nbytes, _stream(dev))
        else:
            fwd = lib.amps_psi_loss_fwd_scan if scan else lib.amps_psi_loss_fwd
            rc = fwd(h, C.byref(p), _ptr(x), B, T, _ptr(loss), _ptr(ws), nbytes,
                     1 if need_grad else 0, _stream(dev))
        _lib.check(h, rc)
        if need_grad:
            ctx.keep = (R_ri, freqs, psi0_ri, x, ws, p, h, model, scan, a_dev, K)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        R_ri, freqs, psi0_ri, x, ws, p, h, model, scan, _a_dev, K = ctx.keep
        lib = _lib.load()
        dev = x.device
        B, T = x.shape
        D = freqs.shape[0]
        w = gloss.detach().contiguous().float()
        ng = int(lib.amps_psi_grad_count(D))
        packed = torch.empty(ng, dtype=torch.float32, device=dev)
        if K > 1:
            rc = lib.amps_psi_loss_bwd_k(h, C.byref(p), _ptr(x), B, T, K, _ptr(w), _ptr(ws), ws.numel(),
                                         _ptr(packed), _stream(dev))
        else:
            bwd = lib.amps_psi_loss_bwd_scan if scan else lib.amps_psi_loss_bwd
            rc = bwd(h, C.byref(p), _ptr(x), B, T, _ptr(w), _ptr(ws), ws.numel(), _ptr(packed), _stream(dev))
        _lib.check(h, rc)
        _allreduce_packed(model, packed, h, dev)
        model._last_packed = packed
        # clones: view_as_real's backward needs an even storage offset (odd D breaks a view)
        gR = packed[: 2 * D * D].view(D, D, 2).clone()
        gf = packed[2 * D * D: 2 * D * D + D].clone()
        gp = packed[2 * D * D + D: 2 * D * D + 3 * D].clone().view(D, 2)
        gA = packed[2 * D * D + 3 * D].clone()
        return gR, gf, gp, gA, None, None, None


class _PsiParamsFn(torch.autograd.Function):
    """Raw variables -> (R_eff, freqs_eff, psi_0, regulariser) in one kernel each way
    (amps_psi_params_fwd / _bwd): model.py:36-50, 218-222; train.py:55-60."""

    @staticmethod
    def forward(ctx, Rx, Ry, fraw, px, py, model):
        dev = Rx.device
        lib = _lib.load()
        h = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        D = fraw.shape[0]
        raws = [t.detach().contiguous().float() for t in (Rx, Ry, fraw, px, py)]
        R_ri = torch.empty(D, D, 2, dtype=torch.float32, device=dev)
        f = torch.empty(D, dtype=torch.float32, device=dev)
        p0 = torch.empty(D, 2, dtype=torch.float32, device=dev)
        aux = torch.empty(4, dtype=torch.float32, device=dev)
        sc = (float(model._r_scale), float(model._f_scale), float(model.h_reg), float(model.r_reg))
        rc = lib.amps_psi_params_fwd(h, D, *[_ptr(t) for t in raws], *sc, _ptr(R_ri), _ptr(f), _ptr(p0),
                                     _ptr(aux), _stream(dev))
        _lib.check(h, rc)
        ctx.keep = (raws, aux, sc, h, D)
        ctx.mark_non_differentiable(aux)
        return R_ri, f, p0, aux[0], aux

    @staticmethod
    def backward(ctx, gR, gf, gp, greg, _gaux):
        raws, aux, sc, h, D = ctx.keep
        lib = _lib.load()
        dev = aux.device
        gR = gR.contiguous().float()
        gf = gf.contiguous().float()
        gp = gp.contiguous().float()
        greg = greg.reshape(1).contiguous().float()
        outs = [torch.empty_like(t) for t in raws]
        rc = lib.amps_psi_params_bwd(h, D, *[_ptr(t) for t in raws], *sc, _ptr(aux), _ptr(gR), _ptr(gf),
                                     _ptr(gp), _ptr(greg), *[_ptr(t) for t in outs], _stream(dev))
        _lib.check(h, rc)
        return (*outs, None)


class _RhoLossFn(torch.autograd.Function):
    """Per-clip RhoCMPS loss through amps_rho_loss_fwd / amps_rho_loss_bwd."""

    @staticmethod
    def forward(ctx, R_ri, freqs, rho0_ri, A, x, model):
        dev = x.device
        if dev.type != "cuda":
            raise RuntimeError("RhoCMPS scan requires a CUDA device (no CPU fallback)")
        lib = _lib.load()
        h = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        B, T = x.shape
        D = freqs.shape[0]
        need_grad = any(ctx.needs_input_grad[:4])
        R_ri = R_ri.detach().contiguous().float()
        freqs = freqs.detach().contiguous().float()
        rho0_ri = rho0_ri.detach().contiguous().float()
        p = _lib.AmpsParams(D=D, reserved=0, R_dev=R_ri.data_ptr(), freqs_dev=freqs.data_ptr(),
                            psi0_dev=None, rho0_dev=rho0_ri.data_ptr(), A=float(A.detach()),
                            sigma=float(model.sigma), delta_t=float(model.delta_t))
        nbytes = lib.amps_rho_workspace_bytes(D, B, T, 1 if need_grad else 0)
        if nbytes == 0:
            raise _lib.AmpsError(-2, f"bond dimension {D} is not supported by the Rho kernels")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        rc = lib.amps_rho_loss_fwd(h, C.byref(p), _ptr(x), B, T, _ptr(loss), _ptr(ws), nbytes,
                                   1 if need_grad else 0, _stream(dev))
        _lib.check(h, rc)
        if need_grad:
            ctx.keep = (R_ri, freqs, rho0_ri, x, ws, p, h, model)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        R_ri, freqs, rho0_ri, x, ws, p, h, model = ctx.keep
        lib = _lib.load()
        dev = x.device
        B, T = x.shape
        D = freqs.shape[0]
        w = gloss.detach().contiguous().float()
        packed = torch.empty(int(lib.amps_rho_grad_count(D)), dtype=torch.float32, device=dev)
        rc = lib.amps_rho_loss_bwd(h, C.byref(p), _ptr(x), B, T, _ptr(w), _ptr(ws), ws.numel(),
                                   _ptr(packed), _stream(dev))
        _lib.check(h, rc)
        _allreduce_packed(model, packed, h, dev)
        model._last_packed = packed
        n = 2 * D * D
        gR = packed[:n].view(D, D, 2).clone()
        gf = packed[n:n + D].clone()
        gr0 = packed[n + D:2 * n + D].clone().view(D, D, 2)
        gA = packed[2 * n + D].clone()
        return gR, gf, gr0, gA, None, None


# --------------------------------------------------------------------------------------------
class CMPS(torch.nn.Module):
    """Continuous Matrix Product State (model.py:5-52)."""

    def __init__(self, hparams, data_iterator=None, freqs_in=None, R_in=None, device=None, seed=0):
        super().__init__()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        self.device = torch.device(device)
        self.hparams = hparams
        self.bond_d = hparams.bond_dim                       # model.py:11
        self.batch_size = hparams.minibatch_size             # :12
        self.h_reg = hparams.h_reg                           # :13
        self.r_reg = hparams.r_reg                           # :14
        self.delta_t = hparams.delta_t                       # :15
        self.dt = np.float32(hparams.delta_t)                # :16
        self.sigma = hparams.sigma                           # :21
        self.data_iterator = data_iterator                   # :25
        self._dp_group = None
        self._dp_native = False
        self._last_packed = None
        gen = torch.Generator(device="cpu").manual_seed(seed)
        self._gen = gen
        D = self.bond_d

        def P(v):
            return torch.nn.Parameter(torch.as_tensor(np.asarray(v), dtype=torch.float32).to(self.device))

        self.A = P(np.float32(hparams.A))                    # :19
        if R_in is not None:                                 # :31-33
            R_in = np.asarray(R_in)
            self.Rx = P(np.ascontiguousarray(R_in.real))
            self.Ry = P(np.ascontiguousarray(R_in.imag))
            self._r_scale = 1.0
        else:                                                # :36-39
            self.Rx = P(torch.randn(D, D, generator=gen).numpy())
            self.Ry = P(torch.randn(D, D, generator=gen).numpy())
            self._r_scale = _rsqrt32(self.r_reg)
        if freqs_in is not None:                             # :44-46
            self.freqs_raw = P(np.asarray(freqs_in))
            self._f_scale = 1.0
        else:                                                # :49-50
            self.freqs_raw = P(torch.randn(D, generator=gen).numpy())
            self._f_scale = _rsqrt32(self.h_reg)

    # TF variable names under scope "model/" (train.py:49; SURVEY 5, checkpoint row)
    TF_NAMES = {"A": "A", "Rx": "Rx", "Ry": "Ry", "freqs_raw": "freqs", "psi_x": "psi_x",
                "psi_y": "psi_y", "Wx": "Wx", "Wy": "Wy"}

    # ---- TensorFlow V2 checkpoints by the reference's variable names -------------
    def save_tf_checkpoint(self, prefix: str, global_step: int = 0) -> None:
        from . import tf_checkpoint as tfc
        tfc.write_tf_checkpoint(prefix, tfc.model_to_tf_variables(self, global_step))

    def load_tf_checkpoint(self, prefix_or_dir: str) -> int:
        """Load `model/Rx, Ry, freqs, psi_x, psi_y | Wx, Wy, A` from a TF checkpoint prefix, or from
        the latest checkpoint of a directory (its `checkpoint` state file); returns global_step."""
        import os
        from . import tf_checkpoint as tfc
        prefix = prefix_or_dir
        if os.path.isdir(prefix_or_dir):
            prefix = tfc.latest_checkpoint(prefix_or_dir)
            if prefix is None:
                raise FileNotFoundError(f"no `checkpoint` state file in {prefix_or_dir}")
        return tfc.load_tf_variables(self, tfc.read_tf_checkpoint(prefix))

    @property
    def R(self) -> torch.Tensor:
        """Effective complex R (model.py:41-42): R[i,j] - R[j,j] (the broadcast quirk)."""
        R = torch.complex(self._r_scale * self.Rx, self._r_scale * self.Ry)
        return R - torch.diagonal(R)

    @property
    def freqs(self) -> torch.Tensor:
        return self._f_scale * self.freqs_raw                # model.py:46 / 49-50

    @property
    def freqsc(self) -> torch.Tensor:
        return self.freqs.to(torch.complex64)                # model.py:52

    # ---- helpers -----------------------------------------------------------------------
    def _require_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("the CMPS scan runs only on a CUDA device; there is no CPU fallback")
        _lib.load()

    def _ctx(self):
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        return _lib.context(idx)

    def _batch(self, data=None) -> torch.Tensor:
        if data is None:
            data = self.data_iterator
        if data is None:
            raise ValueError("no data: pass `data` or construct the model with data_iterator=")
        if callable(data):
            data = data()
        elif not isinstance(data, (torch.Tensor, np.ndarray)) and hasattr(data, "__next__"):
            data = next(data)
        x = _as_device_f32(data, self.device)
        if x.dim() != 2:
            raise ValueError(f"data must be [batch, time], got shape {tuple(x.shape)}")
        return x

    def _phases(self, t) -> torch.Tensor:
        t32 = torch.tensor(np.float32(t), dtype=torch.float32, device=self.device)
        ang = self.freqs * t32
        return torch.complex(torch.cos(ang), torch.sin(ang))  # exp(1j*freqsc*t), model.py:305

    def _noise(self, num_samples, length, temp, noise, generator):
        if noise is None:
            g = generator if generator is not None else self._gen
            std = float(self.sigma * np.sqrt(temp * self.delta_t))          # model.py:246
            noise = torch.randn(length, num_samples, generator=g, dtype=torch.float32) * std
        noise = _as_device_f32(noise, self.device)
        if noise.shape != (length, num_samples):
            raise ValueError(f"noise must be [length, num_samples]={length, num_samples}, got {tuple(noise.shape)}")
        return noise

    def set_data_parallel(self, group, native: bool = False):
        """All-reduce (sum) the packed kernel gradient over ``group`` inside backward.

        ``native=True``: the reduction runs on the library's own NCCL communicator
        (amps_comm_init / amps_allreduce_grads); ``group`` is then used once, to hand rank 0's
        ncclUniqueId to the other ranks."""
        import torch.distributed as dist
        if group is None and dist.is_available() and dist.is_initialized():
            group = dist.group.WORLD          # "default group" is represented explicitly, never as None
        self._dp_group = group
        self._dp_native = False
        if native:
            lib, h = _lib.load(), self._ctx()
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            buf = (C.c_ubyte * 128)()
            if rank == 0:
                _lib.check(h, lib.amps_comm_unique_id(buf))
            ids = [bytes(buf)]
            dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            raw = (C.c_ubyte * 128).from_buffer_copy(ids[0])
            _lib.check(h, lib.amps_comm_init(h, raw, rank, world))
            self._dp_native = True


# --------------------------------------------------------------------------------------------
class PsiCMPS(CMPS):
    """Evolves the state (model.py:206-334)."""

    def __init__(self, hparams, psi_in=None, *args, **kwargs):
        super().__init__(hparams, *args, **kwargs)
        D = self.bond_d
        if psi_in is not None:
            # the reference's psi_in branch is broken (model.py:214-216 read undefined names);
            # the evident intent -- initialise from a complex vector -- is what is done here.
            psi_in = np.asarray(psi_in)
            px, py = np.ascontiguousarray(psi_in.real), np.ascontiguousarray(psi_in.imag)
        else:
            lim = math.sqrt(3.0 / D)  # TF default glorot-uniform on shape [D] (model.py:218-219)
            px = ((torch.rand(D, generator=self._gen) * 2 - 1) * lim).numpy()
            py = ((torch.rand(D, generator=self._gen) * 2 - 1) * lim).numpy()
        self.psi_x = torch.nn.Parameter(torch.as_tensor(px, dtype=torch.float32).to(self.device))
        self.psi_y = torch.nn.Parameter(torch.as_tensor(py, dtype=torch.float32).to(self.device))

    @property
    def psi_0(self) -> torch.Tensor:
        return self._normalize_psi(torch.complex(self.psi_x, self.psi_y))   # model.py:221-222

    # ---- public ------------------------------------------------------------------------
    #: "auto" | "never" | "always": when ``loss_per_clip`` takes the parallel-in-time tensor-core scan
    #: instead of the one-chain-per-clip kernels.  Both meet the parity tolerances; "auto" picks the
    #: scan where it measured faster on B200 (profiles/r1_scan_timings.md): few long clips, D <= 64.
    time_parallel = "auto"

    #: checkpoint interval K of the training path (amps_psi_loss_fwd_k / _bwd_k): 1 keeps the whole
    #: state trajectory for the adjoint sweep (fastest, 12 + 16 D bytes per step and clip), an int K > 1
    #: keeps one state per K steps and recomputes window by window in the backward, "auto" keeps the
    #: trajectory while it fits in ``checkpoint_auto_fraction`` of the free device memory and switches
    #: to K = 2048 beyond that (the reference's own O(T) activation stack: model.py:265-266).
    checkpoint_every = "auto"
    checkpoint_auto_fraction = 0.5

    def _checkpoint_interval(self, D: int, B: int, T: int) -> int:
        K = self.checkpoint_every
        if K is None:
            return 1
        if K == "auto":
            # decided once per shape: querying the free memory every step costs a driver call per forward
            key = (D, B, T, self.checkpoint_auto_fraction)
            cache = self.__dict__.setdefault("_ckpt_auto_cache", {})
            if key not in cache:
                cache[key] = self._checkpoint_auto(D, B, T)
            return cache[key]
        K = int(K)
        if K < 1:
            raise ValueError(f"checkpoint_every must be >= 1, 'auto' or None (got {K})")
        return K

    def _checkpoint_auto(self, D: int, B: int, T: int) -> int:
        full = _lib.load().amps_psi_workspace_bytes(D, B, T, 1)
        free, _total = torch.cuda.mem_get_info(self.device)
        # memory the caching allocator holds but has not handed out is reusable too
        free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        return 1 if full <= self.checkpoint_auto_fraction * free else 2048

    def _use_scan(self, B: int, T: int, need_grad: bool) -> bool:
        if self.time_parallel == "always":
            return True
        if self.time_parallel != "auto" or self.bond_d > 64 or T < 4096 or B < 1:
            return False
        if self.bond_d > 32:
            return B <= (16 if need_grad else 8)
        return B <= 4

    def loss_per_clip(self, data=None, time_parallel: Optional[bool] = None) -> torch.Tensor:
        """loss_b of the fold, before the reduce_mean (model.py:257-267); differentiable.
        ``time_parallel``: force (True) or forbid (False) the parallel-in-time scan; None = policy."""
        return self.loss_per_clip_and_regulariser(data, time_parallel)[0]

    def loss_per_clip_and_regulariser(self, data=None, time_parallel: Optional[bool] = None):
        """(loss_b [B], h_reg |freqs|^2 + r_reg |R|^2): the raw -> effective parameter chain of
        model.py:36-50, 218-222 and the regulariser of train.py:55-60 come from one fused kernel
        (amps_psi_params_fwd) instead of a few dozen framework launches; both are differentiable."""
        self._require_cuda()
        x = self._batch(data)
        if time_parallel is None:
            need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
            time_parallel = self._use_scan(x.shape[0], x.shape[1], need_grad)
        R_ri, f, p0_ri, reg, _ = _PsiParamsFn.apply(self.Rx, self.Ry, self.freqs_raw, self.psi_x, self.psi_y, self)
        lpc = _PsiLossFn.apply(R_ri, f, p0_ri, self.A, x, self, bool(time_parallel))
        return lpc, reg

    def loss_fn(self, data=None) -> torch.Tensor:
        return self.loss_per_clip(data).mean()                              # model.py:267

    @property
    def loss(self) -> torch.Tensor:
        """``model.loss`` of the reference (train.py:59,71), evaluated on data_iterator."""
        return self.loss_fn()

    def grads(self, data=None, weights=None) -> dict:
        """d(sum_b w_b loss_b)/d raw variables, w_b = 1/B by default."""
        lpc = self.loss_per_clip(data)
        tot = lpc.mean() if weights is None else (lpc * _as_device_f32(weights, self.device)).sum()
        names = [n for n, _ in self.named_parameters()]
        gs = torch.autograd.grad(tot, [p for _, p in self.named_parameters()], allow_unused=True)
        return {self.TF_NAMES.get(n, n): g for n, g in zip(names, gs)}

    def loss_per_clip_scan(self, data=None) -> torch.Tensor:
        """``loss_per_clip`` forced onto the parallel-in-time tensor-core scan (amps_psi_loss_fwd_scan /
        amps_psi_loss_bwd_scan): same values and gradients, D <= 64; differentiable."""
        return self.loss_per_clip(data, time_parallel=True)

    def sample(self, num_samples, length, temp=1, noise=None, generator=None) -> torch.Tensor:
        """[num_samples, length] cumulative X_t scaled by A (model.py:242-251)."""
        return self.sample_from_noise(self._noise(num_samples, length, temp, noise, generator))

    @torch.no_grad()
    def sample_from_noise(self, noise) -> torch.Tensor:
        self._require_cuda()
        noise = _as_device_f32(noise, self.device)
        L, n = noise.shape
        lib, h = _lib.load(), self._ctx()
        R_ri = torch.view_as_real(self.R).contiguous()
        f = self.freqs.contiguous()
        p0 = torch.view_as_real(self.psi_0).contiguous()
        p = _lib.AmpsParams(D=self.bond_d, reserved=0, R_dev=R_ri.data_ptr(), freqs_dev=f.data_ptr(),
                            psi0_dev=p0.data_ptr(), rho0_dev=None, A=float(self.A),
                            sigma=float(self.sigma), delta_t=float(self.delta_t))
        out = torch.empty(n, L, dtype=torch.float32, device=self.device)
        nbytes = lib.amps_psi_sample_workspace_bytes(self.bond_d, L, n)
        if nbytes == 0:
            raise _lib.AmpsError(-2, f"bond dimension {self.bond_d} is not supported by the Psi sampler")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        rc = lib.amps_psi_sample(h, C.byref(p), _ptr(noise), L, n, _ptr(out), _ptr(ws), nbytes,
                                 _stream(self.device))
        _lib.check(h, rc)
        return out

    @torch.no_grad()
    def psi_evolve_with_data(self, data=None) -> torch.Tensor:
        """complex64 [B, T-1, D]: normalised psi after every step (model.py:231-240)."""
        self._require_cuda()
        x = self._batch(data)
        B, T = x.shape
        D = self.bond_d
        lib, h = _lib.load(), self._ctx()
        R_ri = torch.view_as_real(self.R).contiguous()
        f = self.freqs.contiguous()
        p0 = torch.view_as_real(self.psi_0).contiguous()
        p = _lib.AmpsParams(D=D, reserved=0, R_dev=R_ri.data_ptr(), freqs_dev=f.data_ptr(),
                            psi0_dev=p0.data_ptr(), rho0_dev=None, A=float(self.A),
                            sigma=float(self.sigma), delta_t=float(self.delta_t))
        nbytes = lib.amps_psi_workspace_bytes(D, B, T, 1)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        out = torch.empty(B, max(T - 1, 0), D, 2, dtype=torch.float32, device=self.device)
        rc = lib.amps_psi_evolve(h, C.byref(p), _ptr(x), B, T, _ptr(out), _ptr(ws), nbytes,
                                 _stream(self.device))
        _lib.check(h, rc)
        return torch.view_as_complex(out)

    # ---- private single-step primitives (tested by the reference, tests/test_model.py:124-138) --
    def _update_ancilla_psi(self, psi, signal, t):
        """One un-normalised update (model.py:300-317); single-step host math, not the scan."""
        psi = torch.as_tensor(psi).to(self.device, torch.complex64)
        signal = (_as_device_f32(signal, self.device) / self.A).to(torch.complex64)
        phases = self._phases(t)
        Upsi = psi * torch.conj(phases)
        R = self.R
        RUpsi = torch.einsum("bc,ac->ab", R, Upsi)
        RdagRUpsi = torch.einsum("bc,ac->ab", R.conj().transpose(0, 1), RUpsi)
        delta_Upsi = -self.delta_t * self.sigma ** 2 * RdagRUpsi / 2.0
        delta_Upsi = delta_Upsi + signal.unsqueeze(1) * RUpsi
        return psi + phases * delta_Upsi

    def _expectation(self, psi, t):
        psi = torch.as_tensor(psi).to(self.device, torch.complex64)
        Upsi = psi * torch.conj(self._phases(t))
        return 2 * torch.einsum("ab,bc,ac->a", torch.conj(Upsi), self.R, Upsi).real

    def _normalize_psi(self, x, axis=None, epsilon=1e-12):
        sq = torch.square(torch.abs(x))
        ssum = sq.sum() if axis is None else sq.sum(dim=axis, keepdim=True)
        return x * torch.rsqrt(torch.clamp(ssum, min=epsilon)).to(x.dtype)


# --------------------------------------------------------------------------------------------
class RhoCMPS(CMPS):
    """Evolves the density matrix (model.py:55-203)."""

    def __init__(self, hparams, W_in=None, *args, **kwargs):
        super().__init__(hparams, *args, **kwargs)
        D = self.bond_d
        rank = getattr(hparams, "initial_rank", None)
        self.rank_rho_0 = rank if rank is not None else D                   # model.py:62-65
        if W_in is not None:
            W_in = np.asarray(W_in)
            wx, wy = np.ascontiguousarray(W_in.real), np.ascontiguousarray(W_in.imag)
        else:
            lim = math.sqrt(6.0 / (self.rank_rho_0 + D))                    # glorot-uniform
            wx = ((torch.rand(self.rank_rho_0, D, generator=self._gen) * 2 - 1) * lim).numpy()
            wy = ((torch.rand(self.rank_rho_0, D, generator=self._gen) * 2 - 1) * lim).numpy()
        self.Wx = torch.nn.Parameter(torch.as_tensor(wx, dtype=torch.float32).to(self.device))
        self.Wy = torch.nn.Parameter(torch.as_tensor(wy, dtype=torch.float32).to(self.device))

    @property
    def rho_0(self) -> torch.Tensor:
        W = torch.complex(self.Wx, self.Wy)                                  # model.py:127
        rho_0 = W.conj().transpose(0, 1) @ W                                 # :128
        return rho_0 / torch.einsum("ii->", rho_0)                           # :129

    def _params(self, keep):
        R_ri = torch.view_as_real(self.R.detach()).contiguous()
        f = self.freqs.detach().contiguous()
        r0 = torch.view_as_real(self.rho_0.detach()).contiguous()
        keep.extend([R_ri, f, r0])
        return _lib.AmpsParams(D=self.bond_d, reserved=0, R_dev=R_ri.data_ptr(),
                               freqs_dev=f.data_ptr(), psi0_dev=None, rho0_dev=r0.data_ptr(),
                               A=float(self.A), sigma=float(self.sigma), delta_t=float(self.delta_t))

    def loss_per_clip(self, data=None) -> torch.Tensor:
        """loss_b of the rho fold, before the reduce_mean (model.py:132-142); differentiable."""
        self._require_cuda()
        x = self._batch(data)
        return _RhoLossFn.apply(torch.view_as_real(self.R), self.freqs,
                                torch.view_as_real(self.rho_0), self.A, x, self)

    def loss_fn(self, data=None):
        return self.loss_per_clip(data).mean()                               # model.py:142

    @property
    def loss(self):
        return self.loss_fn()

    @torch.no_grad()
    def rho_evolve_with_data(self, data=None) -> torch.Tensor:
        """complex64 [B, T-1, D, D] (model.py:76-85)."""
        self._require_cuda()
        x = self._batch(data)
        B, T = x.shape
        D = self.bond_d
        lib, h, keep = _lib.load(), self._ctx(), []
        p = self._params(keep)
        nbytes = lib.amps_rho_workspace_bytes(D, B, T, 0)
        if nbytes == 0:
            raise _lib.AmpsError(-2, f"bond dimension {D} is not supported by the Rho kernels")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, max(T - 1, 0), D, D, 2, dtype=torch.float32, device=self.device)
        rc = lib.amps_rho_evolve(h, C.byref(p), _ptr(x), B, T, _ptr(out), _ptr(ws), nbytes,
                                 _stream(self.device))
        _lib.check(h, rc)
        return torch.view_as_complex(out)

    @torch.no_grad()
    def _sample_scan(self, noise, want_out, want_traj, want_purity):
        self._require_cuda()
        noise = _as_device_f32(noise, self.device)
        L, n = noise.shape
        D = self.bond_d
        lib, h, keep = _lib.load(), self._ctx(), []
        p = self._params(keep)
        nbytes = lib.amps_rho_workspace_bytes(D, n, L + 1, 0)
        if nbytes == 0:
            raise _lib.AmpsError(-2, f"bond dimension {D} is not supported by the Rho kernels")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        out = torch.empty(n, L, dtype=torch.float32, device=self.device) if want_out else None
        traj = torch.empty(n, L, D, D, 2, dtype=torch.float32, device=self.device) if want_traj else None
        pur = torch.empty(n, L, dtype=torch.float32, device=self.device) if want_purity else None
        rc = lib.amps_rho_sample(h, C.byref(p), _ptr(noise), L, n, _ptr(out), _ptr(traj), _ptr(pur),
                                 _ptr(ws), nbytes, _stream(self.device))
        _lib.check(h, rc)
        return out, (torch.view_as_complex(traj) if traj is not None else None), pur

    def sample(self, num_samples, length, temp=1, noise=None, generator=None):
        """[num_samples, length] (model.py:103-112)."""
        noise = self._noise(num_samples, length, temp, noise, generator)
        return self._sample_scan(noise, True, False, False)[0]

    def sample_from_noise(self, noise):
        return self._sample_scan(noise, True, False, False)[0]

    def rho_evolve_with_sampling(self, num_samples, length, temp=1, noise=None, generator=None):
        """complex64 [num_samples, length, D, D] (model.py:87-93)."""
        noise = self._noise(num_samples, length, temp, noise, generator)
        return self._sample_scan(noise, False, True, False)[1]

    def purity(self, num_samples, length, temp=1, noise=None, generator=None):
        """Re tr(rho_k^2), [num_samples, length] (model.py:95-101)."""
        noise = self._noise(num_samples, length, temp, noise, generator)
        return self._sample_scan(noise, False, False, True)[2]

    # ---- private single-step primitives (tests/test_model.py:69-83) ------------------------
    def _Rt(self, t):
        phases = self._phases(t)
        return torch.einsum("a,ab,b->ab", phases, self.R, torch.conj(phases))   # model.py:179

    def _update_ancilla_rho(self, rho, signal, t):
        """One un-normalised update U rho U^dag (model.py:172-187); single-step host math."""
        rho = torch.as_tensor(rho).to(self.device, torch.complex64)
        signal = (_as_device_f32(signal, self.device) / self.A).to(torch.complex64)
        batch = rho.shape[0]
        Rt = self._Rt(t)
        RR_dag = (Rt.conj().transpose(0, 1) @ Rt).unsqueeze(0)
        IR = torch.einsum("a,bc->abc", signal, Rt)
        one = torch.eye(self.bond_d, dtype=torch.complex64, device=self.device).unsqueeze(0).repeat(batch, 1, 1)
        U = one + (-0.5 * RR_dag * self.delta_t * self.sigma ** 2 + IR)
        return torch.einsum("abc,acd,ade->abe", U, rho, U.conj().transpose(1, 2))

    def _expectation(self, rho, t):
        Rt = self._Rt(t)
        x = Rt + Rt.conj().transpose(0, 1)
        return torch.einsum("ab,cba->c", x, torch.as_tensor(rho).to(self.device, torch.complex64)).real

    def _normalize_rho(self, x, epsilon=1e-12):
        tr = torch.einsum("aii->a", x).reshape(-1, 1, 1)
        return x * torch.reciprocal(torch.clamp(tr.real, min=epsilon)).to(x.dtype)
