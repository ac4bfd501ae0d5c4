"""Command-line trainer with the flags of the reference's ``train.py`` (train.py:18-33, 36-94):

    python -m audio_mps_b200.train_cli --mps_model=psi_mps --dataset=damped_sine \\
        --sample_duration=65536 --hparams=bond_dim=32,minibatch_size=64 --logdir=/tmp/amps --steps=100

Launch under ``torchrun`` for batch data parallelism (one process per GPU; the batch is sharded and the
packed kernel gradient is all-reduced once per step).  The summaries of train.py:62-85 go to TensorBoard event
files in ``{logdir}`` -- scalars every step (also as JSON lines in ``{logdir}/scalars.jsonl``); every
``--save_summaries_steps`` steps the ``data`` audio clips (5 at most), the ``frequencies`` histogram and, with
``--visualize``, the ``data_waveform`` / ``sample_waveform`` images (train.py:74-85).  A checkpoint
(``model.pt``: raw variables under the reference's names, Adam state, global_step) is written every
``--save_checkpoint_secs`` (60 s in the reference, train.py:93) and restored on restart.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import HParams, PsiCMPS, RhoCMPS, default_hparams, get_audio
from .train import Trainer, regulariser, shard_bounds


def waveform_image(waveform, width=300, height=300):
    """The picture of ``utils.waveform_plot`` (utils.py:10-17: signal against time, 3 x 3 inch figure) without
    matplotlib: a [3, H, W] uint8 raster, white background, axes box, the waveform as a polyline."""
    w = np.asarray(waveform, np.float64).ravel()
    img = np.full((height, width), 255, np.uint8)
    m = 12                                                           # margin of the axes box
    img[m, m:width - m] = img[height - m - 1, m:width - m] = 0
    img[m:height - m, m] = img[m:height - m, width - m - 1] = 0
    if w.size:
        lo, hi = float(w.min()), float(w.max())
        if hi - lo < 1e-30:
            lo, hi = lo - 1.0, hi + 1.0
        nx = width - 2 * m - 2
        # min / max of the samples that fall into each pixel column (a polyline through every sample)
        edges = np.linspace(0, w.size, nx + 1).astype(np.int64)
        for c in range(nx):
            seg = w[edges[c]:max(edges[c + 1], edges[c] + 1)]
            if seg.size == 0:
                continue
            y0 = (hi - float(seg.max())) / (hi - lo) * (height - 2 * m - 3)
            y1 = (hi - float(seg.min())) / (hi - lo) * (height - 2 * m - 3)
            img[m + 1 + int(y0):m + 2 + int(y1), m + 1 + c] = 40
    return np.stack([img, img, img])


def write_summaries(tb, step, batch, model, sample_rate, visualize=False, samples=None):
    """train.py:74-85: audio summary of the batch (max_outputs = 5), histogram of the frequencies in Hz, and (with
    --visualize) waveform images of the batch and of samples drawn from the model."""
    batch = np.asarray(batch, np.float32)
    for i in range(min(5, batch.shape[0])):
        tb.add_audio(f"data/{i}", torch.from_numpy(batch[i:i + 1]), step, sample_rate=sample_rate)
    with torch.no_grad():
        tb.add_histogram("frequencies", (model.freqs / (2 * math.pi)).detach().float().cpu(), step)
    if visualize:
        for i in range(batch.shape[0]):
            tb.add_image(f"data_waveform/{i}", waveform_image(batch[i]), step)
        if samples is not None:
            for i, w in enumerate(np.asarray(samples)):
                tb.add_image(f"sample_waveform/{i}", waveform_image(w), step)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--mps_model", default="psi_mps", choices=["rho_mps", "psi_mps"])                # train.py:18
    ap.add_argument("--dataset", default="damped_sine", choices=["damped_sine", "guitar", "organ", "nsynth"])
    ap.add_argument("--sample_duration", type=int, default=2 ** 16)                                  # train.py:27
    ap.add_argument("--sample_rate", type=int, default=16000)
    ap.add_argument("--num_samples", type=int, default=3)
    ap.add_argument("--hparams", default="")
    ap.add_argument("--datadir", default="./data")
    ap.add_argument("--logdir", default="../logging/audio_mps")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--save_checkpoint_secs", type=float, default=60.0)
    ap.add_argument("--save_summaries_steps", type=int, default=100,
                    help="cadence of the audio / histogram / image summaries (tf.contrib.training.train default)")
    ap.add_argument("--visualize", action="store_true", help="waveform images of data and samples (train.py:24, 77-85)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--init_tf_checkpoint", default="",
                    help="TensorFlow V2 checkpoint prefix or directory (the reference's logdir) to start from")
    ap.add_argument("--save_tf_checkpoint", action="store_true",
                    help="also write model.ckpt-<step> in TensorFlow V2 format at the end of the run")
    args = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    hp = default_hparams(args.sample_rate)
    hp.parse(args.hparams)                                                                           # train.py:44
    cls = RhoCMPS if args.mps_model == "rho_mps" else PsiCMPS
    model = cls(hp, device=dev, seed=args.seed)                     # same seed on every rank: replicas
    trainer = Trainer(model)
    logdir = os.path.join(args.logdir, args.dataset, f"{hp.bond_dim}_{hp.delta_t}_{hp.minibatch_size}")  # train.py:94
    ckpt = os.path.join(logdir, "model.pt")
    if rank == 0:
        os.makedirs(logdir, exist_ok=True)
    if os.path.exists(ckpt):                                        # MonitoredTrainingSession-style restore
        trainer.load_state_dict(torch.load(ckpt, map_location=dev))
    elif args.init_tf_checkpoint:                                   # weights trained by the reference
        trainer.global_step = model.load_tf_checkpoint(args.init_tf_checkpoint)
    rng = np.random.default_rng(args.seed + 1)
    source = get_audio(args.datadir, args.dataset, hp, sample_duration=args.sample_duration, rng=rng)
    gb = hp.minibatch_size
    lo, hi = shard_bounds(gb, rank, world)
    last_save = time.time()
    log = open(os.path.join(logdir, "scalars.jsonl"), "a") if rank == 0 else None
    tb = None
    if rank == 0:
        try:                                                        # the scalars of train.py:62-72 as event files
            from torch.utils.tensorboard import SummaryWriter
            tb = SummaryWriter(logdir)
        except Exception:
            tb = None

    def save():
        tmp = ckpt + ".tmp"                                         # atomic: never leave a torn model.pt behind
        torch.save(trainer.state_dict(), tmp)
        os.replace(tmp, ckpt)
        if args.save_tf_checkpoint:
            model.save_tf_checkpoint(os.path.join(logdir, f"model.ckpt-{trainer.global_step}"), trainer.global_step)
    for _ in range(args.steps):
        batch = source if isinstance(source, np.ndarray) else next(source)
        if isinstance(source, np.ndarray):                          # a fresh random-onset batch per step
            source = get_audio(args.datadir, args.dataset, hp, sample_duration=args.sample_duration, rng=rng)
        model_loss = float(trainer.step(batch[lo:hi], global_batch=batch.shape[0]))
        if rank == 0:
            with torch.no_grad():
                R, f = model.R, model.freqs
                h_l2 = float(torch.sum(f * f))
                r_l2 = float(torch.sum(torch.conj(R) * R).real)
                # total_loss pairs the step's model loss with the regulariser of the SAME (pre-update) parameters
                reg = float(trainer.last_reg) if trainer.last_reg is not None else hp.h_reg * h_l2 + hp.r_reg * r_l2
                rec = {"step": trainer.global_step, "model_loss": model_loss,
                       "total_loss": model_loss + reg,
                       "A": float(model.A), "sigma": float(model.sigma),
                       "h_l2norm": math.sqrt(h_l2), "r_l2norm": math.sqrt(r_l2),
                       "gr_decay_time": 1.0 / (2 * math.pi * hp.sigma ** 2 * r_l2 / hp.bond_dim)}   # train.py:66-67
            log.write(json.dumps(rec) + "\n")
            log.flush()
            if tb is not None:
                for k, v in rec.items():
                    if k != "step":
                        tb.add_scalar(k, v, trainer.global_step)
                if args.save_summaries_steps > 0 and (trainer.global_step - 1) % args.save_summaries_steps == 0:
                    smp = None
                    if args.visualize and args.num_samples:
                        smp = model.sample(args.num_samples, args.sample_duration).cpu().numpy()
                    write_summaries(tb, trainer.global_step, batch, model, args.sample_rate, args.visualize, smp)
            if time.time() - last_save >= args.save_checkpoint_secs:
                save()
                last_save = time.time()
    if rank == 0:
        save()
        if tb is not None:
            tb.close()
        if args.num_samples:
            w = model.sample(args.num_samples, args.sample_duration)                                 # train.py:83
            np.save(os.path.join(logdir, "samples.npy"), w.cpu().numpy())
        log.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
