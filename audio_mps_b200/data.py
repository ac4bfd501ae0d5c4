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
