"""TensorFlow-free reader / writer for TensorFlow "V2" checkpoints (tensor bundles), so that weights
trained by the reference (`tf.train.MonitoredTrainingSession(checkpoint_dir=...)`, train.py:88-93)
can be evaluated on this path and vice versa.  SURVEY 8(f4).

Variables the reference saves (model.py:19,32-50,122-126,215-219 under `tf.variable_scope("model")`,
train.py:49; `tf.train.get_or_create_global_step()`, train.py:88):
    model/A, model/Rx, model/Ry, model/freqs, model/psi_x, model/psi_y (Psi) or model/Wx, model/Wy
    (Rho), global_step -- plus Adam's slots model/<v>/Adam, model/<v>/Adam_1, beta1_power, beta2_power.

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table = the LevelDB table format):
  <prefix>.index                 sorted string table: "" -> BundleHeaderProto, name -> BundleEntryProto
  <prefix>.data-00000-of-00001   raw little-endian tensor bytes at (offset, size) of the entry
  checkpoint                     text proto naming the latest prefix (CheckpointState)
No TensorFlow is installable in this environment, so this module is pinned only by its own
round trip and by the published format (parity against a real TF-written file is unpinned).
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Optional

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 8: np.dtype("<c8")}
_DT_INV = {v: k for k, v in _DT.items()}

# ---- crc32c (Castagnoli), masked as LevelDB / TensorFlow store it -------------------------------
_CRC_TABLE = []
for _n in range(256):
    _c = _n
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    tab = _CRC_TABLE
    c = 0xFFFFFFFF
    for byte in data:
        c = tab[(c ^ byte) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / minimal protobuf ------------------------------------------------------------------
def _put_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf: bytes, pos: int):
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _pb_fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _pb_varint(num: int, v: int) -> bytes:
    return _put_varint(num << 3) + _put_varint(v)


def _pb_bytes(num: int, v: bytes) -> bytes:
    return _put_varint((num << 3) | 2) + _put_varint(len(v)) + v


def _entry_proto(dtype: np.dtype, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)       # TensorShapeProto.dim
    out = _pb_varint(1, _DT_INV[dtype]) + _pb_bytes(2, dims)
    if offset:
        out += _pb_varint(4, offset)
    out += _pb_varint(5, size)
    out += _put_varint((6 << 3) | 5) + struct.pack("<I", crc)
    return out


def _parse_entry(buf: bytes):
    dtype, shape, shard, offset, size, crc = 0, [], 0, 0, 0, None
    for num, wt, v in _pb_fields(buf):
        if num == 1:
            dtype = v
        elif num == 2:
            for n2, _, dim in _pb_fields(v):
                if n2 == 2:
                    sz = 0
                    for n3, _, dv in _pb_fields(dim):
                        if n3 == 1:
                            sz = dv if dv < (1 << 63) else dv - (1 << 64)
                    shape.append(sz)
        elif num == 3:
            shard = v
        elif num == 4:
            offset = v
        elif num == 5:
            size = v
        elif num == 6:
            crc = struct.unpack("<I", v)[0]
        elif num == 7:
            raise ValueError("partitioned (sliced) variables are not supported")
    return dtype, tuple(shape), shard, offset, size, crc


# ---- snappy (decompression only; TF writes the index uncompressed, accepted for robustness) ------
def _snappy_decompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block")
    return bytes(out)


# ---- table blocks -----------------------------------------------------------------------------------
def _read_block(f: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    raw = f[offset:offset + size]
    ctype = f[offset + size]
    if verify:
        stored = struct.unpack("<I", f[offset + size + 1:offset + size + 5])[0]
        if stored != masked_crc32c(f[offset:offset + size + 1]):
            raise ValueError("table block checksum mismatch")
    if ctype == 0:
        return raw
    if ctype == 1:
        return _snappy_decompress(raw)
    raise ValueError(f"unknown table block compression {ctype}")


def _block_entries(block: bytes):
    nrestarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * nrestarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _build_block(entries, restart_interval: int = 16) -> bytes:
    out = bytearray()
    restarts = []
    prev = b""
    for n, (k, v) in enumerate(entries):
        if n % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _emit_block(f: bytearray, block: bytes):
    off = len(f)
    f += block + b"\x00"
    f += struct.pack("<I", masked_crc32c(block + b"\x00"))
    return _put_varint(off) + _put_varint(len(block))


# ---- public API --------------------------------------------------------------------------------------
def read_tf_checkpoint(prefix: str, verify: bool = True) -> Dict[str, np.ndarray]:
    """All tensors of the V2 checkpoint `<prefix>.index` / `<prefix>.data-*` by variable name."""
    idx = open(prefix + ".index", "rb").read()
    if len(idx) < 48 or struct.unpack("<Q", idx[-8:])[0] != _MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorFlow V2 checkpoint index (bad table magic)")
    footer = idx[-48:]
    _, p = _get_varint(footer, 0)                 # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isz, p = _get_varint(footer, p)
    entries = {}
    for _, handle in _block_entries(_read_block(idx, ioff, isz, verify)):
        boff, q = _get_varint(handle, 0)
        bsz, q = _get_varint(handle, q)
        for k, v in _block_entries(_read_block(idx, boff, bsz, verify)):
            entries[k] = v
    num_shards = 1
    if b"" in entries:
        for num, _, v in _pb_fields(entries[b""]):
            if num == 1:
                num_shards = v
            if num == 2 and v != 0:
                raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    for k, v in entries.items():
        if k == b"":
            continue
        dtype, shape, shard, offset, size, crc = _parse_entry(v)
        if dtype not in _DT:
            raise ValueError(f"{k.decode()}: unsupported dtype enum {dtype}")
        if shard not in shards:
            shards[shard] = open(f"{prefix}.data-{shard:05d}-of-{num_shards:05d}", "rb").read()
        raw = shards[shard][offset:offset + size]
        if verify and crc is not None and masked_crc32c(raw) != crc:
            raise ValueError(f"{k.decode()}: tensor checksum mismatch")
        out[k.decode()] = np.frombuffer(raw, dtype=_DT[dtype]).reshape(shape).copy()
    return out


def write_tf_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], update_state: bool = True) -> None:
    """Write `tensors` as a one-shard V2 checkpoint and (optionally) the `checkpoint` state file."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data = bytearray()
    items = []
    for name in sorted(tensors, key=lambda s: s.encode()):
        a = np.asarray(tensors[name])
        dt = {"f": "<f4", "i": "<i8", "c": "<c8"}.get(a.dtype.kind)
        if a.dtype in (np.float64, np.int32):
            dt = a.dtype.newbyteorder("<").str
        if dt is None:
            raise ValueError(f"{name}: dtype {a.dtype} not supported")
        a = np.asarray(a.astype(np.dtype(dt)), order="C")      # (ascontiguousarray would make scalars 1-d)
        raw = a.tobytes()
        items.append((name.encode(), _entry_proto(np.dtype(dt), a.shape, len(data), len(raw), masked_crc32c(raw))))
        data += raw
    header = _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))        # num_shards = 1, version.producer = 1
    entries = [(b"", header)] + items
    f = bytearray()
    dhandle = _emit_block(f, _build_block(entries))
    mhandle = _emit_block(f, _build_block([]))
    ihandle = _emit_block(f, _build_block([(entries[-1][0] + b"\x00", dhandle)], restart_interval=1))
    footer = mhandle + ihandle
    f += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(f))
    if update_state:
        base = os.path.basename(prefix)
        with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as fh:
            fh.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')


def latest_checkpoint(checkpoint_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint: the prefix named by `<dir>/checkpoint` (CheckpointState)."""
    path = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.exists(path):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(path).read(), flags=re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)


# ---- model glue -----------------------------------------------------------------------------------------
def model_to_tf_variables(model, global_step: int = 0, scope: str = "model") -> Dict[str, np.ndarray]:
    """Raw variables under the reference's names (`model/Rx`, ...), plus `global_step` (int64)."""
    out = {}
    for n, p in model.named_parameters():
        out[f"{scope}/{model.TF_NAMES.get(n, n)}"] = p.detach().cpu().numpy()
    out["global_step"] = np.asarray(global_step, dtype=np.int64)
    return out


def load_tf_variables(model, tensors: Dict[str, np.ndarray], scope: str = "model") -> int:
    """Copy `model/<name>` tensors into the model's raw variables (shapes must match); returns
    the checkpoint's global_step (0 if absent).  Adam slots, if present, are ignored."""
    import torch
    with torch.no_grad():
        for n, p in model.named_parameters():
            key = f"{scope}/{model.TF_NAMES.get(n, n)}"
            if key not in tensors:
                raise KeyError(f"checkpoint has no variable {key!r} (has: {sorted(tensors)[:12]} ...)")
            a = np.asarray(tensors[key], dtype=np.float32)
            if tuple(a.shape) != tuple(p.shape):
                raise ValueError(f"{key}: checkpoint shape {a.shape} != model shape {tuple(p.shape)}")
            p.copy_(torch.from_numpy(a).to(p.device))
    return int(tensors.get("global_step", 0))
