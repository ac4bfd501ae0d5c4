"""Input side of the path: the reference's synthetic clip generator (/root/reference/data.py:8-22)
and a TensorFlow-free reader for its ``{audio: float32[sample_duration]}`` TFRecords
(data.py:27-43).  Returns float32 ``[batch, sample_duration]`` batches."""
from __future__ import annotations

import struct
from typing import Iterator, Optional

import numpy as np


def damped_sine(batch: int, length: int, delta_t: float, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """261.6 Hz sine, 0.1 s decay, onset delay ~ Gamma(2, rate 2/(length/100)) (data.py:10-20)."""
    rng = rng or np.random.default_rng(0)
    freq = 261.6
    decay_time = 0.1
    delay_time = length / 100
    delays = rng.gamma(shape=2.0, scale=delay_time / 2.0, size=(batch, 1)).astype(np.float32)
    input_range = np.arange(length, dtype=np.float32)[None, :]
    times = ((input_range - delays) * np.float32(delta_t)).astype(np.float32)
    wave = 0.5 * (np.sign(times) + 1) * np.sin(2 * np.pi * freq * times) * np.exp(-times / decay_time)
    return wave.astype(np.float32)


def random_raw_params(bond_dim: int, A: float, rng: np.random.Generator) -> dict:
    """Raw Psi trainables drawn with the reference's initialisers (model.py:36-39, 49-50, 218-219:
    standard normal Rx, Ry, freqs; glorot-uniform psi_x, psi_y on a [D] vector), in the draw order the
    synthetic benchmark / golden inputs are defined with (default_rng(seed))."""
    D = int(bond_dim)
    lim = float(np.sqrt(3.0 / D))
    return {"Rx": rng.standard_normal((D, D)).astype(np.float32),
            "Ry": rng.standard_normal((D, D)).astype(np.float32),
            "freqs": rng.standard_normal(D).astype(np.float32),
            "A": np.float32(A),
            "psi_x": rng.uniform(-lim, lim, D).astype(np.float32),
            "psi_y": rng.uniform(-lim, lim, D).astype(np.float32)}


def sample_noise(sigma: float, delta_t: float, length: int, num_samples: int, seed: int, temp: float = 1.0) -> np.ndarray:
    """The sampler's noise tensor [length, num_samples] ~ N(0, sigma^2 * temp * delta_t), drawn once
    (model.py:246), from a recorded seed."""
    z = np.random.default_rng(seed).standard_normal((length, num_samples)).astype(np.float32)
    return z * np.float32(sigma * np.sqrt(temp * delta_t))


# ---- TFRecord / tf.train.Example parsing without TensorFlow --------------------------------
def _read_varint(buf: bytes, pos: int):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) of one protobuf message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def parse_example_float_feature(record: bytes, key: str) -> np.ndarray:
    """Extract a FloatList feature from a serialised tf.train.Example.

    Example{1: Features{1: map<string, Feature>}}; Feature{2: FloatList{1: packed floats}}."""
    for fn, _, features in _fields(record):
        if fn != 1:
            continue
        for fn2, _, entry in _fields(features):
            if fn2 != 1:
                continue
            name, feat = None, None
            for fn3, _, v in _fields(entry):
                if fn3 == 1:
                    name = v.decode()
                elif fn3 == 2:
                    feat = v
            if name != key or feat is None:
                continue
            for fn4, _, fl in _fields(feat):
                if fn4 == 2:  # float_list
                    vals = []
                    for fn5, wt5, pv in _fields(fl):
                        if fn5 == 1 and wt5 == 2:
                            vals.append(np.frombuffer(pv, dtype="<f4"))
                        elif fn5 == 1 and wt5 == 5:
                            vals.append(np.frombuffer(pv, dtype="<f4"))
                    return np.concatenate(vals) if vals else np.zeros(0, np.float32)
    raise KeyError(f"feature {key!r} not found in record")


def iter_tfrecords(path: str) -> Iterator[bytes]:
    """TFRecord framing: u64 length, u32 crc, payload, u32 crc (crcs not verified)."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if len(head) < 12:
                return
            (length,) = struct.unpack("<Q", head[:8])
            payload = f.read(length)
            f.read(4)
            if len(payload) < length:
                raise ValueError("truncated TFRecord")
            yield payload


def write_tfrecords(path: str, clips: np.ndarray, key: str = "audio") -> None:
    """Write ``{key: float32[...]}`` Examples (make-small-dataset.py:18-34); crc fields zeroed."""
    def varint(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)

    def ld(fn, payload):
        return varint((fn << 3) | 2) + varint(len(payload)) + payload

    with open(path, "wb") as f:
        for clip in np.asarray(clips, dtype="<f4"):
            float_list = ld(1, clip.tobytes())
            feature = ld(2, float_list)
            entry = ld(1, key.encode()) + ld(2, feature)
            features = ld(1, entry)
            example = ld(1, features)
            f.write(struct.pack("<Q", len(example)) + b"\0\0\0\0" + example + b"\0\0\0\0")


def tfrecord_batches(path: str, batch_size: int, sample_duration: int, key: str = "audio",
                     repeat: bool = True) -> Iterator[np.ndarray]:
    """batch -> (shuffle 24, omitted: order is file order) -> repeat (data.py:38-41)."""
    while True:
        buf = []
        for rec in iter_tfrecords(path):
            a = parse_example_float_feature(rec, key)
            if a.shape[0] != sample_duration:
                raise ValueError(f"record has {a.shape[0]} samples, expected {sample_duration}")
            buf.append(a)
            if len(buf) == batch_size:
                yield np.stack(buf).astype(np.float32)
                buf = []
        if buf:
            yield np.stack(buf).astype(np.float32)
        if not repeat:
            return


def get_audio(datadir, dataset, hps, sample_duration: int = 2 ** 16, rng=None):
    """Same call as the reference's ``get_audio`` (data.py:6): a float32 [B, sample_duration]
    array for 'damped_sine', else an iterator of such batches from ``{datadir}/{dataset}.tfrecords``."""
    if dataset == "damped_sine":
        return damped_sine(hps.minibatch_size, sample_duration, hps.delta_t, rng)
    return tfrecord_batches(f"{datadir}/{dataset}.tfrecords", hps.minibatch_size, sample_duration)


class DeviceBatchPrefetcher:
    """Double-buffered host -> device staging of [B, T] float32 batches for the training loop: the
    copy of the next batch (pinned host memory, its own CUDA stream) overlaps the scan kernels of
    the current step; the compute stream only waits on the copy's event.

        pf = DeviceBatchPrefetcher(device, (B, T))
        pf.submit(first_host_batch)
        for ...:
            x = pf.next()                # device tensor, ready on the current stream
            pf.submit(next_host_batch)   # starts copying while the step below runs
            trainer.step(x)
    """

    def __init__(self, device, shape):
        import torch
        self._torch = torch
        self.device = device
        self._bufs = [torch.empty(shape, dtype=torch.float32, device=device) for _ in range(2)]
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._free = [torch.cuda.Event() for _ in range(2)]
        self._stream = torch.cuda.Stream(device)
        self._head = 0      # next buffer to fill
        self._tail = 0      # next buffer to hand out
        self._pending = 0

    def submit(self, host_batch) -> None:
        torch = self._torch
        if self._pending >= 2:
            raise RuntimeError("both staging buffers are in flight: call next() first")
        if not host_batch.is_pinned():
            host_batch = host_batch.pin_memory()
        i = self._head
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(self._free[i])          # the step that used this buffer is done
            self._bufs[i].copy_(host_batch, non_blocking=True)
            self._ready[i].record(self._stream)
        self._head ^= 1
        self._pending += 1

    def next(self):
        torch = self._torch
        if self._pending == 0:
            raise RuntimeError("no batch submitted")
        i = self._tail
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[i])
        self._tail ^= 1
        self._pending -= 1
        self._last = i
        return self._bufs[i]

    def release(self) -> None:
        """Mark the batch handed out by the last next() as consumed (call after the step is enqueued)."""
        self._free[self._last].record(self._torch.cuda.current_stream(self.device))
