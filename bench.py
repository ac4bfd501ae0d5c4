#!/usr/bin/env python
"""bench.py -- AudioMPS fwd+bwd audio samples/s on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
  python bench.py --impl reference --gpus N ...           # the reference's CPU algorithm (oracle port)
  python bench.py --config c3|c4 ...                      # the other training configs of BASELINE.json

A "step" is one training step of the hot path over one batch of synthetic clips: per-clip loss
(forward scan), adjoint backward, raw-parameter chain + regulariser (train.py:55-60) and Adam
(train.py:89).  Workload at every N: BASELINE.json configs[1] PER GPU (D=32, 64 clips of 4 s at
16 kHz = 64000 samples), i.e. weak scaling: global batch = 64*N, clips sharded over ranks, one
NCCL all-reduce of the packed gradient per step.

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

METRIC = "AudioMPS fwd+bwd audio samples/s (D=32, 4 s 16 kHz clips)"
UNIT = "samples/s"
D, B_PER_GPU, T = 32, 64, 64000
WORKLOAD = "C1: PsiCMPS training step D=32, 64 clips/GPU x 64000 samples (4 s @ 16 kHz), damped-sine clips"
# --config: the other training configs of BASELINE.json through the same harness (default: C1, the
# configuration the metric is quoted on)
CONFIGS = {
    "c1": (32, 64, "C1"),
    "c3": (128, 128, "C3 (bond dimension and batch; run on the row-split cluster kernels)"),
    "c4": (64, 256, "C4 (per-GPU share of the global batch 2048 at 8 GPUs)"),
}


def select_config(name):
    global D, B_PER_GPU, WORKLOAD, METRIC
    D, B_PER_GPU, tag = CONFIGS[name]
    if name != "c1":
        METRIC = f"AudioMPS fwd+bwd audio samples/s (D={D}, 4 s 16 kHz clips)"
        WORKLOAD = (f"{tag}: PsiCMPS training step D={D}, {B_PER_GPU} clips/GPU x {T} samples "
                    f"(4 s @ 16 kHz), damped-sine clips")


def hparams_kw():
    return dict(minibatch_size=B_PER_GPU, bond_dim=D, delta_t=1 / 16000, sigma=0.0001,
                h_reg=200 / (np.pi * 16000) ** 2, r_reg=0.1, initial_rank=None, A=100.,
                learning_rate=0.001)


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
def cpu_port_step(t_sample, seed=1):
    """One fwd+bwd of the op-for-op PyTorch-CPU restatement (oracle) on B_PER_GPU clips x t_sample
    time steps of the C1 workload.  Returns seconds."""
    import torch
    from oracle.cmps_oracle import HP, PsiCMPSOracle, damped_sine, grads_of, random_raw_params, total_loss
    hp = HP(**hparams_kw())
    raw = random_raw_params(hp, np.random.default_rng(0))
    full = damped_sine(B_PER_GPU, T, hp.delta_t, np.random.default_rng(seed))
    data = np.ascontiguousarray(full[:, 4000:4000 + t_sample + 1])   # inside the sounding part of the clips
    t0 = time.perf_counter()
    m = PsiCMPSOracle(hp, raw, mode="f32")
    loss = total_loss(m, data)
    grads_of(m, loss)
    return time.perf_counter() - t0


def c_port_rate(t_sample):
    """The compiled C/OpenMP restatement (oracle/cmps_ref.c), float32 arithmetic, all host threads."""
    try:
        from oracle import cref
        hp_kw = hparams_kw()
        return cref.bench_loss_grad(D, B_PER_GPU, t_sample, hp_kw)
    except Exception as e:  # the C port is optional
        return None


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (TensorFlow itself cannot be installed here;
    DESIGN.md) = the oracle's op-for-op PyTorch-CPU float32 restatement, all host threads."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t_sample = max(100, int(args.ref_tsample * (32 / D) ** 2 * 64 / B_PER_GPU))
    for _ in range(args.warmup):
        cpu_port_step(min(t_sample, 200))
    times = [cpu_port_step(t_sample) for _ in range(args.steps)]
    sec = float(np.mean(times))
    val = B_PER_GPU * t_sample / sec
    sample = (f"{B_PER_GPU} clips x {t_sample} consecutive samples of the C1 clips (D={D}), "
              f"fwd+bwd via torch autograd, PyTorch-CPU complex64 op-for-op port of model.py")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "bounded_sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from audio_mps_b200 import DeviceBatchPrefetcher, HParams, PsiCMPS, _lib, damped_sine
    from audio_mps_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    hp = HParams(**hparams_kw())
    model = PsiCMPS(hp, device=dev, seed=0)            # same seed on every rank: replicated params
    trainer = Trainer(model, group=None)
    gb = B_PER_GPU * world
    x_host = torch.from_numpy(damped_sine(B_PER_GPU, T, hp.delta_t, np.random.default_rng(1 + rank))).pin_memory()
    x_dev = x_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def events():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    for _ in range(max(args.warmup, 3)):
        trainer.step(x_dev, global_batch=gb)
    barrier()

    # ---- timed region: ONE barrier + synchronize bracket around the K steps (no host sync inside:
    # A travels by device pointer, so the host queues ahead); per-step CUDA event pairs leave the L2
    # flush between iterations out of the sum.
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count(local)
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations
        e0, e1 = events()
        e0.record()
        trainer.step(x_dev, global_batch=gb)
        e1.record()
        evs.append((e0, e1))
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    launches = _lib.launch_count(local) - launches0

    # ---- kernel durations (library event pairs on the launch stream), a few extra synchronised steps
    _lib.set_profiling(local, True)
    fwd_ms, bwd_ms = [], []
    for _ in range(min(args.steps, 3)):
        flush.zero_()
        trainer.step(x_dev, global_batch=gb)
        barrier()
        fwd_ms.append(_lib.kernel_ms(local, 0))
        bwd_ms.append(_lib.kernel_ms(local, 1))
    _lib.set_profiling(local, False)

    # ---- end to end: every step's batch comes from pinned host memory (double-buffered prefetch on
    # a copy stream, DeviceBatchPrefetcher) and its loss goes back to pinned host memory, all inside
    # the bracket
    pf = DeviceBatchPrefetcher(dev, (B_PER_GPU, T))

    def e2e_run(k):
        marks = []
        pf.submit(x_host)
        for i in range(k):
            flush.zero_()
            e0, e1 = events()
            e0.record()
            x = pf.next()
            if i + 1 < k:
                pf.submit(x_host)                       # next step's H2D overlaps this step's kernels
            ml = trainer.step(x, global_batch=gb)
            pf.release()
            loss_host.copy_(ml.reshape(1), non_blocking=True)   # D2H of the step's result
            e1.record()
            marks.append((e0, e1))
        return marks
    e2e_run(2)
    barrier()
    e0a, e1a = events()
    e0a.record()
    marks = e2e_run(args.steps)
    e1a.record()
    barrier()
    # the first batch's copy is not overlapped: count the whole bracket minus the flushes' share
    e2e_ms = [a.elapsed_time(b) for a, b in marks]
    e2e_ms[0] = max(e2e_ms[0], e0a.elapsed_time(marks[0][1]))
    final_loss = float(loss_host[0])
    clocks = sampler.stop()       # sampled from the start of the timed region to the end of the e2e arm

    tot = torch.tensor([sum(step_ms), sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot_ms, tot_e2e_ms = float(tot[0]), float(tot[1])

    if rank == 0:
        K = args.steps
        samples = gb * T * K
        value = samples / (tot_ms * 1e-3)
        e2e_val = samples / (tot_e2e_ms * 1e-3)
        peaks, peak_src = measured_peaks()
        # Algorithmic traffic per (clip, sample), DESIGN.md 4.6 (K = 1: the whole trajectory is kept):
        #   forward  4 B waveform + 8*D B x_k + 8*D B S x'_k + 8 B (E_k, |x_k|^2) written
        #   backward the same bytes read back
        bwd_t = float(np.mean(bwd_ms)) * 1e-3
        fwd_t = float(np.mean(fwd_ms)) * 1e-3
        units = B_PER_GPU * T
        per_unit_bytes = 12 + 16 * D
        # executed flops: fwd = L_k formation 4D^2 + chain mat-vec 8D^2 + S x' 8D^2 ;
        #                 bwd = L_k^dag formation 4D^2 + chain mat-vec 8D^2 + 3 rank-1 tiles 24D^2
        fwd_flops = units * (20 * D * D + 40 * D)
        bwd_flops = units * (36 * D * D + 60 * D)
        fma_peak = float(_lib.load().amps_fma_peak_tflops(_lib.context(local)))
        clustered = D <= 32 and 2 * B_PER_GPU <= 148 and os.environ.get("AMPS_NO_CLUSTER") != "1"
        kn = ("cl_kernel<%d,4>" % D) if clustered else ("kernel<%d,4>" % D) if D <= 32 else \
            "uni_kernel<64,8>" if D <= 64 else "c4_kernel<128,4>"
        kinfo = {"fwd": (f"psi_fwd_{kn}", fwd_t, fwd_flops), "bwd": (f"psi_bwd_{kn}", bwd_t, bwd_flops)}
        dom = "fwd" if fwd_t >= bwd_t else "bwd"
        oth = "bwd" if dom == "fwd" else "fwd"

        def entry(which):
            name, tt, fl = kinfo[which]
            return {"kernel": name, "kernel_ms": tt * 1e3, "achieved": units * per_unit_bytes / tt / 1e9,
                    "frac": units * per_unit_bytes / tt / 1e9 / peaks["hbm_gbs"],
                    "fp32_achieved_tflops": fl / tt / 1e12,
                    "fp32_frac": fl / tt / 1e12 / fma_peak if fma_peak > 0 else None,
                    "cycles_per_step_at_1965MHz": tt / (T - 1) * 1.965e9}
        d = entry(dom)
        # dram__bytes_read+write per launch of the same kernels at this exact workload, from the
        # committed ncu --set full capture (profiles/r1_ncu_cluster.md)
        ncu_traffic = {"fwd": 2.106e9, "bwd": 2.172e9} if (clustered and D == 32 and B_PER_GPU == 64) \
            else {"fwd": None, "bwd": None}
        roof = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved"], "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": d["frac"], "traffic": ncu_traffic[dom],
                "traffic_source": "profiles/r1_ncu_cluster.md", "peak_source": peak_src,
                "kernel_ms": d["kernel_ms"],
                "note": ("the path is dependent-step-latency / FP32-issue bound, not HBM bound (SURVEY 0.10): "
                         "one clip per SM pair advances one step per ~300-360 cycles; see fp32 and other_kernel")
                if D <= 32 else ("not HBM bound: the 16-warp D >= 64 kernels are bound by the LSU data pipe "
                                 "(profiles/r1_ncu_uni.md); see fp32 and other_kernel"),
                "fp32": {"achieved_tflops": d["fp32_achieved_tflops"], "peak_tflops": fma_peak,
                         "peak_source": "FFMA microbenchmark in this run (amps_fma_peak_tflops)",
                         "frac": d["fp32_frac"]},
                "cycles_per_step": d["cycles_per_step_at_1965MHz"],
                "other_kernel": entry(oth)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            ts = max(100, int(args.ref_tsample * (32 / D) ** 2 * 64 / B_PER_GPU))
            cpu_port_step(100)
            sec = cpu_port_step(ts)
            cpu = {"value": B_PER_GPU * ts / sec, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{B_PER_GPU} clips x {ts} consecutive samples of the C1 clips, fwd+bwd, "
                             f"PyTorch-CPU complex64 op-for-op port of model.py (TensorFlow not installable)"}
            crate = c_port_rate(ts * 4)
            if crate is not None:
                cpu["c_openmp_port_value"] = crate
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": gb, "bond_dim": D, "samples_per_clip": T,
                           "parallelism": f"dp{world}", "l2": "256 MiB flush between timed steps",
                           "step": "fwd scan + adjoint bwd + regulariser + Adam"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B_PER_GPU * T * 4,
                        "d2h_bytes_per_step": 4, "ms_per_step": tot_e2e_ms / K},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "final_loss": final_loss}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-tsample", type=int, default=1500,
                    help="time steps per CPU-baseline step (bounded sample of the 64000-sample clips)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c1", choices=sorted(CONFIGS),
                    help="BASELINE.json training config (default c1 = the one the metric is quoted on)")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
