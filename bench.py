#!/usr/bin/env python
"""bench.py -- AudioMPS fwd+bwd audio samples/s on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path (default config c1)
  python bench.py --impl reference --gpus N ...           # the reference's CPU algorithm (C/OpenMP float32 port)
  python bench.py --config c0|c1|c2|c3|c4 ...             # any BASELINE.json config through the same harness

A "step" is one pass of the hot path over one batch of synthetic input:
  training configs (c0, c1, c3, c4): per-clip loss (forward scan), adjoint backward, raw-parameter chain
      + regulariser (train.py:55-60) and Adam (train.py:89);
  sampling config (c2): PsiCMPS.sample of 256 waveforms x 64000 steps from a fixed noise tensor.
Workload at every N: BASELINE.json configs[1] PER GPU (D=32, 64 clips of 4 s at 16 kHz), i.e. weak
scaling: global batch = 64*N, clips sharded over ranks, one NCCL all-reduce of the packed gradient per
step.  At N = 1 the default run also measures the other four BASELINE configs (`other_configs`) with
their parity error against the committed goldens, and at N > 1 it checks that the replicas stayed
identical (`dp_check`).

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "samples/s"
T_FULL = 64000
# name -> (kind, D, batch per GPU, T, tag, golden fixture[, checkpoint interval K])
CONFIGS = {
    "c0": ("train", 8, 8, 16000, "C0: reference CPU-runnable case", "psi_c0_d8_t16000"),
    "c1": ("train", 32, 64, T_FULL, "C1", "psi_c1_full"),
    "c2": ("sample", 32, 256, T_FULL, "C2: sampling from a fixed noise tensor", "psi_c2_sample_full"),
    "c3": ("train", 128, 128, T_FULL, "C3 (bond dimension and batch; row-split 4-CTA cluster kernels)",
           "psi_c3_batch_t2000"),
    "c4": ("train", 64, 256, T_FULL, "C4 (per-GPU share of the global batch 2048 at 8 GPUs)",
           "psi_c4_batch_t4000"),
    # checkpointed backward (one state per K steps, windowed recompute): C1 at K = 2048, and C4's WHOLE global
    # batch on one GPU (136 GB of trajectory at K = 1)
    "c1_k2048": ("train", 32, 64, T_FULL, "C1 with checkpoint interval K = 2048", None, 2048),
    "c4_b2048_k2048": ("train", 64, 2048, T_FULL, "C4's global batch 2048 on ONE GPU, checkpoint interval K = 2048",
                       None, 2048),
}


class Cfg:
    def __init__(self, name):
        self.name = name
        self.kind, self.D, self.B, self.T, self.tag, self.golden = CONFIGS[name][:6]
        self.K = CONFIGS[name][6] if len(CONFIGS[name]) > 6 else None
        if self.kind == "sample":
            self.metric = f"AudioMPS sampling audio samples/s (D={self.D}, 4 s 16 kHz clips)"
            self.workload = (f"{self.tag}: PsiCMPS.sample D={self.D}, {self.B} waveforms/GPU x {self.T} steps, "
                             f"noise [L, n] ~ N(0, sigma^2 dt) drawn once")
        else:
            self.metric = f"AudioMPS fwd+bwd audio samples/s (D={self.D}, 4 s 16 kHz clips)"
            self.workload = (f"{self.tag}: PsiCMPS training step D={self.D}, {self.B} clips/GPU x {self.T} samples"
                             f"{' (4 s @ 16 kHz)' if self.T == T_FULL else ' (1 s @ 16 kHz)'}, damped-sine clips")

    def hparams_kw(self):
        return dict(minibatch_size=self.B, bond_dim=self.D, delta_t=1 / 16000, sigma=0.0001,
                    h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100.,
                    learning_rate=0.001)

    # SURVEY 8(d) algorithmic work per (clip, sample)
    def flops(self, which):
        D = self.D
        return {"fwd": 24 * D * D + 36 * D, "bwd": 72 * D * D + 94 * D, "sample": 24 * D * D + 40 * D}[which]

    def kernel_names(self, n_sms=148):
        D, B = self.D, self.B
        if self.kind == "sample":
            return {"sample": f"psi_sample_kernel<{_dp(D)}>"}
        clustered = D <= 32 and 2 * B <= n_sms and os.environ.get("AMPS_NO_CLUSTER") != "1"
        kn = (f"cl_kernel<{_dp(D)}>" if clustered else f"kernel<{_dp(D)}>") if D <= 32 else \
            "uni_kernel<64,4>" if D <= 64 else "c4_kernel<128,4,...,128 threads>"
        names = {"fwd": f"psi_fwd_{kn}", "bwd": f"psi_bwd_{kn}"}
        if D > 32 and os.environ.get("AMPS_NO_TC_TILES") != "1":
            # chain-only adjoint sweep + the gradient tiles as tcgen05 GEMMs over the time axis (both inside "bwd")
            names["bwd"] += " (chain only) + psi_tiles_tc_kernel"
            names["tiles"] = f"psi_tiles_tc_kernel<{_dp(D)}>"
        return names


def _dp(D):
    for p in (8, 16, 32, 64, 128):
        if D <= p:
            return p
    return D


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU arms (the ONLY place bench.py executes oracle/): the reference's algorithm on the host cores
# ---------------------------------------------------------------------------------------------
def _effective_params(cfg):
    """float32 effective parameters of the benchmark model (seed 0), via the oracle's parameter chain."""
    from oracle import cref
    from oracle.cmps_oracle import HP, PsiCMPSOracle, random_raw_params
    hp = HP(**cfg.hparams_kw())
    raw = random_raw_params(hp, np.random.default_rng(0))
    o = PsiCMPSOracle(hp, raw, mode="f32", requires_grad=False)
    return hp, cref.effective_from_oracle(o)


def c_port_eval(cfg, t_steps=None):
    """One evaluation of the compiled C/OpenMP float32 restatement (oracle/cmps_ref.c), all host
    threads: loss + adjoint gradient (training configs) or the sampler (c2).  t_steps=None = the FULL
    configuration.  Returns (seconds, units processed)."""
    from oracle import cref
    from oracle.cmps_oracle import damped_sine
    hp, (R, f, p0, A) = _effective_params(cfg)
    if cfg.kind == "sample":
        from audio_mps_b200.data import sample_noise
        L = cfg.T if t_steps is None else t_steps
        noise = sample_noise(hp.sigma, hp.delta_t, L, cfg.B, 2)
        t0 = time.perf_counter()
        cref.psi_sample(R, f, p0, A, hp.sigma, hp.delta_t, noise, mode="f32")
        return time.perf_counter() - t0, cfg.B * L
    full = damped_sine(cfg.B, cfg.T, hp.delta_t, np.random.default_rng(1))
    if t_steps is None:
        x = full
    else:
        off = min(4000, max(0, cfg.T - t_steps - 1))          # inside the sounding part of the clips
        x = np.ascontiguousarray(full[:, off:off + t_steps + 1])
    t0 = time.perf_counter()
    cref.psi_loss_grad(R, f, p0, A, hp.sigma, hp.delta_t, x, mode="f32")
    return time.perf_counter() - t0, cfg.B * x.shape[1]


def torch_port_step(cfg, t_sample, seed=1):
    """One fwd+bwd of the op-for-op PyTorch-CPU restatement (oracle/cmps_oracle.py) on a bounded sample."""
    from oracle.cmps_oracle import HP, PsiCMPSOracle, damped_sine, grads_of, random_raw_params, total_loss
    hp = HP(**cfg.hparams_kw())
    raw = random_raw_params(hp, np.random.default_rng(0))
    full = damped_sine(cfg.B, cfg.T, hp.delta_t, np.random.default_rng(seed))
    off = min(4000, max(0, cfg.T - t_sample - 1))
    data = np.ascontiguousarray(full[:, off:off + t_sample + 1])
    t0 = time.perf_counter()
    m = PsiCMPSOracle(hp, raw, mode="f32")
    loss = total_loss(m, data)
    grads_of(m, loss)
    return time.perf_counter() - t0


def c_port_full_is_affordable(cfg):
    # full-size evaluations that end in well under a minute on a 16-core host (measured: C1 ~10 s, C2 ~10 s)
    return cfg.name in ("c0", "c1", "c2")


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU algorithm for this path.  TensorFlow 1.x cannot be
    installed here, so it is the oracle's compiled C/OpenMP float32 port, all host threads,
    on the SAME configuration as the CUDA arm where a full evaluation takes seconds (c0, c1, c2); the
    D >= 64 configs (hours of CPU) use a bounded time sample of the same clips."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    full = c_port_full_is_affordable(cfg)
    t_steps = None if full else max(200, int(3000 * (32 / cfg.D) ** 2 * 64 / cfg.B))
    for _ in range(min(args.warmup, 2)):
        c_port_eval(cfg, 500)
    times, units = [], 0
    for _ in range(args.steps):
        sec, units = c_port_eval(cfg, t_steps)
        times.append(sec)
    sec = float(np.mean(times))
    val = units / sec
    sample = (f"FULL config: {cfg.B} x {cfg.T}" if full else
              f"{cfg.B} clips x {t_steps} consecutive samples of the {cfg.name.upper()} clips") + \
        f" (D={cfg.D}), C/OpenMP float32 restatement of model.py (oracle/cmps_ref.c), {cores} threads"
    line = {"impl": "reference", "metric": cfg.metric, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg.workload, "same_config": bool(full),
                       **({} if full else {"bounded_sample": sample})},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if cfg.kind == "train" and not args.no_torch_port:
        try:
            import torch
            torch.set_num_threads(cores)
            ts = min(cfg.T - 1, max(100, int(args.ref_tsample * (32 / cfg.D) ** 2 * 64 / cfg.B)))
            torch_port_step(cfg, 100)
            s2 = torch_port_step(cfg, ts)
            line["cpu_baseline"]["torch_eager_port_value"] = cfg.B * ts / s2
            line["cpu_baseline"]["torch_eager_port_sample"] = f"{cfg.B} clips x {ts} samples, torch-CPU complex64 autograd"
        except Exception as e:  # secondary figure only
            line["cpu_baseline"]["torch_eager_port_error"] = str(e)[:200]
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------
class Harness:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA GPU: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        from audio_mps_b200 import _lib
        self.lib = _lib
        _lib.load()
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.args = args
        self._fma = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def events(self):
        E = self.torch.cuda.Event
        return E(enable_timing=True), E(enable_timing=True)

    def fma_peak(self):
        if self._fma is None:
            self._fma = float(self.lib.load().amps_fma_peak_tflops(self.lib.context(self.local)))
        return self._fma

    def timed(self, step, steps, warmup):
        """W untimed steps, then K steps inside ONE barrier + synchronize bracket; per-step CUDA event
        pairs leave the L2 flush between iterations out of the sum.  Returns (ms list, launches)."""
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        l0 = self.lib.launch_count(self.local)
        evs = []
        self.barrier()
        for _ in range(steps):
            self.flush.zero_()                              # L2 flush between timed iterations
            e0, e1 = self.events()
            e0.record()
            step()
            e1.record()
            evs.append((e0, e1))
        self.barrier()
        return [a.elapsed_time(b) for a, b in evs], self.lib.launch_count(self.local) - l0

    def kernel_ms(self, step, which, reps):
        """durations of the library's dominant kernels (event pairs on the launch stream)"""
        self.lib.set_profiling(self.local, True)
        out = {w: [] for w in which}
        for _ in range(reps):
            self.flush.zero_()
            step()
            self.barrier()
            for w in which:
                out[w].append(self.lib.kernel_ms(self.local, w))
        self.lib.set_profiling(self.local, False)
        return {w: float(np.mean(v)) for w, v in out.items()}


def golden_parity(h, cfg):
    """Worst per-clip relative loss error (training configs) / worst per-path sample error (c2) of
    the CUDA path against the committed golden fixture of this config (tests/golden/, minted by the
    float64 restatement; the parity tests hold the same numbers to 1e-4 / 1e-3)."""
    torch = h.torch
    from audio_mps_b200 import HParams, PsiCMPS, damped_sine, random_raw_params, sample_noise
    path = os.path.join(ROOT, "tests", "golden", cfg.golden + ".npz")
    if not os.path.exists(path):
        return None
    g = dict(np.load(path, allow_pickle=False))
    kw = cfg.hparams_kw()
    try:
        if cfg.name == "c0":
            D, B, T, seed = int(g["hp"][0]), int(g["B"]), int(g["T"]), int(g["seed"])
            raw = {k[4:]: g[k] for k in g if k.startswith("raw_")}
            data = damped_sine(B, T, kw["delta_t"], np.random.default_rng(seed + 1))
        elif cfg.kind == "sample":
            D, B, T, seed = int(g["D"]), int(g["n"]), int(g["L"]), int(g["seed"])
            raw = random_raw_params(D, kw["A"], np.random.default_rng(seed))
        else:
            D, B, T, seed = int(g["D"]), int(g["B"]), int(g["T"]), int(g["seed"])
            raw = random_raw_params(D, kw["A"], np.random.default_rng(seed))
            off = int(g["offset"]) if "offset" in g else 0
            full = damped_sine(B, T_FULL if "offset" in g else T, kw["delta_t"], np.random.default_rng(seed + 1))
            data = np.ascontiguousarray(full[:, off:off + T])
        kw.update(bond_dim=D, minibatch_size=B)
        m = PsiCMPS(HParams(**kw), device=h.dev)
        with torch.no_grad():
            for k, v in raw.items():
                getattr(m, "freqs_raw" if k == "freqs" else k).copy_(torch.as_tensor(np.asarray(v, np.float32)))
            if cfg.kind == "sample":
                stride = int(g["stride"])
                out = m.sample_from_noise(sample_noise(kw["sigma"], kw["delta_t"], T, B, seed + 2))
                sub = out[:, stride - 1::stride].double().cpu().numpy()
                err = float((np.abs(sub - g["sub"]) / g["absmax"][:, None]).max())
                return {"golden": cfg.golden, "what": "worst sample path, |dX| / max|X| (tolerance 1e-3)", "err": err}
            got = m.loss_per_clip(data).double().cpu().numpy()
        ref = g["loss_f64"]
        den = np.maximum(np.abs(ref), 1e-2 * np.abs(ref).max())
        return {"golden": cfg.golden, "what": "worst per-clip loss, relative (tolerance 1e-4)",
                "err": float((np.abs(got - ref) / den).max())}
    except Exception as e:
        return {"golden": cfg.golden, "error": str(e)[:200]}


def measure_train(h, cfg, steps, warmup, full_detail):
    torch = h.torch
    from audio_mps_b200 import DeviceBatchPrefetcher, HParams, PsiCMPS, damped_sine
    from audio_mps_b200.train import Trainer
    D, B, T = cfg.D, cfg.B, cfg.T
    hp = HParams(**cfg.hparams_kw())
    model = PsiCMPS(hp, device=h.dev, seed=0)            # same seed on every rank: replicated params
    if cfg.K is not None:
        model.checkpoint_every = cfg.K
    trainer = Trainer(model, group=None, cuda_graph=os.environ.get("AMPS_BENCH_NO_GRAPH") != "1")
    gb = B * h.world
    x_host = torch.from_numpy(damped_sine(B, T, hp.delta_t, np.random.default_rng(1 + h.rank))).pin_memory()
    x_dev = x_host.to(h.dev)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def step():
        return trainer.step(x_dev, global_batch=gb)

    sampler = ClockSampler(h.local) if full_detail else None
    if sampler:
        sampler.start()
    step_ms, launches = h.timed(step, steps, warmup)
    if trainer.cuda_graph:   # replays launch without passing through the library's host-side counter
        launches = max(launches, steps * trainer.graph_launches_per_step)
    has_tiles = "tiles" in cfg.kernel_names()

    def step_eager():        # per-kernel CUDA events cannot be read out of a captured graph
        return trainer._step_eager(x_dev, global_batch=gb)
    kms = h.kernel_ms(step_eager, (0, 1, 3) if has_tiles else (0, 1), min(steps, 3))
    res = {"step_ms": step_ms, "launches": launches, "fwd_ms": kms[0], "bwd_ms": kms[1], "gb": gb}
    if has_tiles:
        res["tiles_ms"] = kms[3]
    res["checkpoint_interval"] = model._checkpoint_interval(D, B, T)
    res["workspace_bytes"] = int(h.lib.load().amps_psi_workspace_bytes_k(D, B, T, res["checkpoint_interval"]))

    if full_detail:
        # ---- end to end: every step's batch comes from pinned host memory (double-buffered prefetch on
        # a copy stream, DeviceBatchPrefetcher) and its loss goes back to pinned host memory
        pf = DeviceBatchPrefetcher(h.dev, (B, T))

        def e2e_run(k):
            marks = []
            pf.submit(x_host)
            for i in range(k):
                h.flush.zero_()
                e0, e1 = h.events()
                e0.record()
                x = pf.next()
                if i + 1 < k:
                    pf.submit(x_host)                   # next step's H2D overlaps this step's kernels
                ml = trainer.step(x, global_batch=gb)
                pf.release()
                loss_host.copy_(ml.reshape(1), non_blocking=True)   # D2H of the step's result
                e1.record()
                marks.append((e0, e1))
            return marks
        e2e_run(2)
        h.barrier()
        e0a, e1a = h.events()
        e0a.record()
        marks = e2e_run(steps)
        e1a.record()
        h.barrier()
        e2e_ms = [a.elapsed_time(b) for a, b in marks]
        e2e_ms[0] = max(e2e_ms[0], e0a.elapsed_time(marks[0][1]))   # the first copy is not overlapped
        res.update(e2e_ms=e2e_ms, final_loss=float(loss_host[0]), clocks=sampler.stop(),
                   h2d=B * T * 4, d2h=4)
        if h.world > 1:
            res["dp_check"] = dp_check(h, model, trainer, x_dev, gb)
    del trainer, model, x_dev
    torch.cuda.empty_cache()
    return res


def dp_check(h, model, trainer, x_dev, gb):
    """Correctness of the data-parallel step, carried in the bench line because the driver's 1-GPU
    test box skips tests/test_gpu_dp.py: (1) after the K steps every rank holds bit-identical
    parameters; (2) the all-reduced loss of one more step equals the sum of the ranks' local partial
    losses (each rank's sum_b loss_b / global_batch)."""
    torch, dist = h.torch, h.dist
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1).float() for p in model.parameters()])
        hsh = flat.view(torch.int32).to(torch.int64).sum().reshape(1)
        local = (model.loss_per_clip(x_dev).double().sum() / gb).reshape(1)
    hs = [torch.zeros_like(hsh) for _ in range(h.world)]
    dist.all_gather(hs, hsh)
    tot = local.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    glob = trainer.step(x_dev, global_batch=gb).double().reshape(1)
    same = all(int(v) == int(hs[0]) for v in hs)
    rel = float((glob - tot).abs() / tot.abs().clamp_min(1e-30))
    ok = same and rel <= 1e-5
    return {"ok": bool(ok), "params_identical_on_all_ranks": bool(same), "param_hash": int(hs[0]),
            "allreduced_loss": float(glob), "sum_of_rank_losses": float(tot), "rel_diff": rel}


def measure_sample(h, cfg, steps, warmup, full_detail):
    torch = h.torch
    from audio_mps_b200 import HParams, PsiCMPS, sample_noise
    D, n, L = cfg.D, cfg.B, cfg.T
    hp = HParams(**cfg.hparams_kw())
    model = PsiCMPS(hp, device=h.dev, seed=0)
    noise_host = torch.from_numpy(sample_noise(hp.sigma, hp.delta_t, L, n, 2 + h.rank)).pin_memory()
    noise_dev = noise_host.to(h.dev)
    out_host = torch.empty(n, L, dtype=torch.float32).pin_memory()

    def step():
        return model.sample_from_noise(noise_dev)

    sampler = ClockSampler(h.local) if full_detail else None
    if sampler:
        sampler.start()
    step_ms, launches = h.timed(step, steps, warmup)
    kms = h.kernel_ms(step, (2,), min(steps, 3))
    res = {"step_ms": step_ms, "launches": launches, "sample_ms": kms[2], "gb": n * h.world}
    if full_detail:
        def e2e_step():                                   # noise from pinned host memory, waveforms back to it
            out_host.copy_(model.sample_from_noise(noise_host.to(h.dev, non_blocking=True)), non_blocking=True)
        e2e_ms, _ = h.timed(e2e_step, steps, 2)
        res.update(e2e_ms=e2e_ms, final_loss=float(out_host[0, -1]), clocks=sampler.stop(),
                   h2d=n * L * 4, d2h=n * L * 4)
    del model, noise_dev
    torch.cuda.empty_cache()
    return res


def kernel_entries(h, cfg, res):
    """per-kernel roofline figures (SURVEY 8(d) flops and bytes per (clip, sample))"""
    units = cfg.B * cfg.T
    fma = h.fma_peak()
    peaks, peak_src = measured_peaks()
    names = cfg.kernel_names()
    # HBM bytes per unit: SURVEY 8(d) (checkpoint interval K = 64) and what THIS design moves (K = 1: the
    # whole trajectory x_k, S x'_k and (E_k, |x_k|^2) are written by the forward and read by the backward)
    D = cfg.D
    survey_b = {"fwd": 4 + 8 * D / 64, "bwd": 4 + 8 * D / 64, "sample": 8}
    design_b = {"fwd": 12 + 16 * _dp(D), "bwd": 12 + 16 * _dp(D), "sample": 8}
    out = {}
    for which, name in names.items():
        if which == "tiles":
            continue
        ms = res[f"{which}_ms"]
        t = ms * 1e-3
        fl = units * cfg.flops(which)
        out[which] = {"kernel": name, "kernel_ms": ms,
                      "fp32_achieved_tflops": fl / t / 1e12, "fp32_frac": fl / t / 1e12 / fma if fma > 0 else None,
                      "hbm_achieved_gbs_survey": units * survey_b[which] / t / 1e9,
                      "hbm_achieved_gbs_design": units * design_b[which] / t / 1e9,
                      "hbm_frac_design": units * design_b[which] / t / 1e9 / peaks["hbm_gbs"],
                      "bytes_per_unit_survey": survey_b[which], "bytes_per_unit_design": design_b[which],
                      "cycles_per_step_at_1965MHz": t / max(cfg.T - 1, 1) * 1.965e9}
    if "tiles" in names and res.get("tiles_ms"):
        # G_N, G_R, G_E = 3 complex outer products per (clip, sample): 24 D^2 real flops; executed on the tensor
        # pipe as 3 tf32 passes (hi*hi + lo*hi + hi*lo) of the real 2D x 2D (x 4D for the stacked pair) form
        t = res["tiles_ms"] * 1e-3
        Dp = _dp(D)
        out["bwd"]["of_which_tiles_ms"] = res["tiles_ms"]
        out["bwd"]["tiles"] = {"kernel": names["tiles"], "kernel_ms": res["tiles_ms"], "pipe": "tcgen05 kind::tf32, 3-pass split",
                               "algorithmic_tflops": units * 24 * D * D / t / 1e12,
                               "executed_tf32_tflops": units * 3 * 2 * (2 * Dp) * (6 * Dp) / t / 1e12,
                               "hbm_gbs": units * (3 * 8 * Dp + 8) / t / 1e9,
                               "hbm_frac": units * (3 * 8 * Dp + 8) / t / 1e9 / peaks["hbm_gbs"],
                               "bytes_per_unit": 3 * 8 * Dp + 8,
                               "ncu": "profiles/r2_ncu_tiles.md"}
    return out, fma, peaks, peak_src


# dram__bytes_read + dram__bytes_write per launch at the exact bench workload, from the committed
# ncu --set full captures (profiles/): (fwd, bwd) or sampler
NCU_TRAFFIC = {"c1": {"fwd": 2.105e9, "bwd": 2.164e9, "source": "profiles/r2_ncu_summary.md"}}


def run_ours(args, cfg):
    h = Harness(args)
    torch = h.torch
    measure = measure_sample if cfg.kind == "sample" else measure_train
    res = measure(h, cfg, args.steps, args.warmup, True)

    tot = torch.tensor([sum(res["step_ms"]), sum(res["e2e_ms"])], dtype=torch.float64, device=h.dev)
    if h.world > 1:
        h.dist.all_reduce(tot, op=h.dist.ReduceOp.MAX)
    tot_ms, tot_e2e_ms = float(tot[0]), float(tot[1])

    if h.rank == 0:
        K = args.steps
        samples = res["gb"] * cfg.T * K
        value = samples / (tot_ms * 1e-3)
        e2e_val = samples / (tot_e2e_ms * 1e-3)
        ents, fma, peaks, peak_src = kernel_entries(h, cfg, res)
        dom = max(ents, key=lambda k: ents[k]["kernel_ms"])
        d = ents[dom]
        traffic = NCU_TRAFFIC.get(cfg.name, {})
        roof = {"bound": "fp32-latency", "kernel": d["kernel"], "achieved": d["fp32_achieved_tflops"],
                "peak": fma, "unit": "TFLOP/s", "frac": d["fp32_frac"],
                "peak_source": "FFMA issue-rate microbenchmark run in this process (amps_fma_peak_tflops); "
                               "not in MEASURED_PEAKS.json, nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                "flops_per_unit": cfg.flops(dom), "flops_source": "SURVEY 8(d)",
                "traffic": traffic.get(dom), "traffic_source": traffic.get("source"),
                "kernel_ms": d["kernel_ms"], "cycles_per_step": d["cycles_per_step_at_1965MHz"],
                "hbm": {"achieved_gbs_survey_bytes": d["hbm_achieved_gbs_survey"],
                        "achieved_gbs_design_bytes": d["hbm_achieved_gbs_design"],
                        "peak_gbs": peaks["hbm_gbs"], "peak_source": peak_src,
                        "frac_design_bytes": d["hbm_frac_design"],
                        "bytes_per_unit_survey": d["bytes_per_unit_survey"],
                        "bytes_per_unit_design": d["bytes_per_unit_design"]},
                "note": "the path is bound by the dependent-step latency of ONE clip's chain and FP32 issue, "
                        "not by HBM (SURVEY 0.10, 8(d)); the HBM figures are carried because BASELINE asks for them",
                "kernels": ents}
        cpu = None
        other = None
        if h.world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            try:
                c_port_eval(cfg, 300)
                full = c_port_full_is_affordable(cfg)
                ts = None if full else max(200, int(3000 * (32 / cfg.D) ** 2 * 64 / cfg.B))
                sec, units = c_port_eval(cfg, ts)
                cpu = {"value": units / sec, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": ("FULL config, " if full else f"{cfg.B} x {ts} samples of the same clips, ") +
                                 "C/OpenMP float32 restatement of model.py (oracle/cmps_ref.c); TensorFlow 1.x "
                                 "is not installable here"}
            except Exception as e:
                cpu = {"error": str(e)[:200]}
        if h.world == 1 and cfg.name == "c1" and not args.no_other_configs:
            other = {"c1": {"parity_err_vs_golden": golden_parity(h, cfg)}}
            for name in ("c0", "c2", "c3", "c4", "c1_k2048", "c4_b2048_k2048"):
                oc = Cfg(name)
                try:
                    nst = 1 if oc.B > 1024 else 3          # (the 2048-clip step takes ~1.5 s)
                    m2 = (measure_sample if oc.kind == "sample" else measure_train)(h, oc, nst, 3, False)
                    e2, _, _, _ = kernel_entries(h, oc, m2)
                    dk = max(e2, key=lambda k: e2[k]["kernel_ms"])
                    ms = float(np.mean(m2["step_ms"]))
                    other[name] = {"workload": oc.workload, "ms_per_step": ms,
                                   "value": oc.B * oc.T / (ms * 1e-3), "unit": UNIT,
                                   "dominant_kernel": e2[dk]["kernel"], "kernel_ms": e2[dk]["kernel_ms"],
                                   "fp32_frac": e2[dk]["fp32_frac"],
                                   "cycles_per_step": e2[dk]["cycles_per_step_at_1965MHz"],
                                   "kernels_ms": {k: v["kernel_ms"] for k, v in e2.items()},
                                   "gpu_launches": int(m2["launches"]),
                                   "parity_err_vs_golden": golden_parity(h, oc) if oc.golden else None}
                    if "checkpoint_interval" in m2:
                        other[name]["checkpoint_interval"] = m2["checkpoint_interval"]
                        other[name]["workspace_mb"] = m2["workspace_bytes"] / 1e6
                    if "tiles" in e2.get("bwd", {}):
                        other[name]["tiles"] = e2["bwd"]["tiles"]
                except Exception as e:
                    other[name] = {"error": str(e)[:300]}
        line = {"metric": cfg.metric, "value": value, "unit": UNIT, "n_gpus": h.world, "steps": K,
                "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg.workload, "global_batch": res["gb"], "bond_dim": cfg.D,
                           "samples_per_clip": cfg.T, "parallelism": f"dp{h.world}",
                           "l2": "256 MiB flush between timed steps",
                           "step": "sampler" if cfg.kind == "sample" else "fwd scan + adjoint bwd + regulariser + Adam"},
                "clocks": res["clocks"],
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": res["h2d"],
                        "d2h_bytes_per_step": res["d2h"], "ms_per_step": tot_e2e_ms / K},
                "gpu_launches": int(res["launches"]), "roofline": roof, "cpu_baseline": cpu,
                "final_loss": res["final_loss"]}
        if "dp_check" in res:
            line["dp_check"] = res["dp_check"]
        if other is not None:
            line["other_configs"] = other
        print(json.dumps(line))
    if h.world > 1:
        h.dist.barrier()
        h.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-tsample", type=int, default=1500,
                    help="time steps of the secondary torch-eager CPU figure (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the c0/c2/c3/c4 block of the default (c1, N=1) line")
    ap.add_argument("--no-torch-port", action="store_true")
    ap.add_argument("--config", default="c1", choices=sorted(CONFIGS),
                    help="BASELINE.json config (default c1 = the one the metric is quoted on)")
    args = ap.parse_args()
    cfg = Cfg(args.config)
    if args.impl == "reference":
        return run_reference(args, cfg)
    return run_ours(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
