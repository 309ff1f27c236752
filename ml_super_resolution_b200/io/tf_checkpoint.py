"""TensorFlow V2 checkpoints ("tensor bundles": `<prefix>.index` + `<prefix>.data-00000-of-00001`) without TensorFlow
(SURVEY 8f row f3): read the variables a `tf.train.Saver` wrote -- `conv2d/kernel`, `f1/kernel`, `patch_extraction/weights`,
`global_step`, Adam slots -- into the numpy dict the model classes take (`params=`), and write such a bundle back.

Format (tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table, a fork of LevelDB's table format):
  .index  = sorted string table.  Footer (last 48 bytes): BlockHandle(metaindex) BlockHandle(index), each two varint64
            (offset, size), zero padding to 40 bytes, magic 0xdb4775248b80fb57 (little endian).  A block is followed by a
            1-byte compression type (0 = none, 1 = snappy) and a masked crc32c of block + type.  Block contents: entries
            `varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value` (keys are prefix-compressed
            against the previous key; restart points reset `shared` to 0), then `uint32 restarts[n]`, `uint32 n`.
            The index block maps "a key >= the last key of a data block" -> BlockHandle of that data block.
  keys    "" -> BundleHeaderProto {num_shards=1, endianness=2, version=3}; tensor name -> BundleEntryProto
            {dtype=1, shape=2 (TensorShapeProto: repeated Dim dim=2 {size=1}), shard_id=3, offset=4, size=5, crc32c=6 (fixed32)}.
  .data   raw little-endian tensor bytes at [offset, offset + size); crc32c = masked CRC-32C of those bytes.

PARITY UNPINNED: no TensorFlow-written checkpoint is available in this environment; the reader is checked against this
module's own writer and against hand-assembled blocks only.  Snappy-compressed blocks (not what BundleWriter emits) raise.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

from .tfrecord import _fields, _ld, _read_varint, _varint, masked_crc32c

MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"), 6: np.dtype("i1"),
           9: np.dtype("<i8"), 10: np.dtype("?"), 19: np.dtype("<f2")}
_DTYPE_IDS = {v: k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------------- table reader
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    contents = buf[offset:offset + size]
    ctype = buf[offset + size]
    (crc,) = struct.unpack("<I", buf[offset + size + 1:offset + size + 5])
    if verify and crc != masked_crc32c(contents + bytes([ctype])):
        raise IOError("tensor bundle index: block checksum mismatch")
    if ctype != 0:
        raise NotImplementedError("tensor bundle index: compressed block (snappy) -- BundleWriter does not emit these")
    return contents


def _block_entries(block: bytes):
    (n_restarts,) = struct.unpack("<I", block[-4:])
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _handle(buf: bytes, pos: int = 0):
    off, pos = _read_varint(buf, pos)
    size, pos = _read_varint(buf, pos)
    return off, size, pos


def read_index(index_path: str, verify: bool = True) -> "OrderedDict[str, dict]":
    """-> {tensor name: {dtype, shape, shard_id, offset, size, crc32c}} (+ '' -> header fields)."""
    buf = open(index_path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != MAGIC:
        raise IOError(f"{index_path}: not a TensorFlow table file (bad magic)")
    footer = buf[-48:]
    _, _, p = _handle(footer)
    ioff, isize, _ = _handle(footer, p)
    out = OrderedDict()
    for _, hval in _block_entries(_read_block(buf, ioff, isize, verify)):
        doff, dsize, _ = _handle(hval)
        for key, val in _block_entries(_read_block(buf, doff, dsize, verify)):
            name = key.decode()
            if name == "":
                out[""] = {f: v for f, _, v in _fields(val)}
                continue
            e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
            for f, wt, v in _fields(val):
                if f == 1:
                    e["dtype"] = v
                elif f == 2:
                    dims = []
                    for g, _, dv in _fields(v):
                        if g == 2:
                            size = 0
                            for h, _, sv in _fields(dv):
                                if h == 1:
                                    size = sv
                            dims.append(size)
                    e["shape"] = tuple(dims)
                elif f == 3:
                    e["shard_id"] = v
                elif f == 4:
                    e["offset"] = v
                elif f == 5:
                    e["size"] = v
                elif f == 6:
                    e["crc32c"] = struct.unpack("<I", v)[0]
            out[name] = e
    return out


def load_checkpoint(prefix: str, verify: bool = True) -> "OrderedDict[str, np.ndarray]":
    """All tensors of `<prefix>.index` / `<prefix>.data-*` as numpy arrays keyed by variable name; adding ':0' gives the keys
    the model classes use (`{k + ':0': v for k, v in load_checkpoint(p).items()}` or `to_params`)."""
    index = read_index(prefix + ".index", verify)
    n_shards = index.get("", {}).get(1, 1)
    shards = {}
    out = OrderedDict()
    for name, e in index.items():
        if name == "":
            continue
        if e["dtype"] not in _DTYPES:
            raise NotImplementedError(f"{name}: unsupported DataType {e['dtype']}")
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open(f"{prefix}.data-{sid:05d}-of-{n_shards:05d}", "rb").read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise IOError(f"{name}: data shard truncated")
        if verify and e["crc32c"] is not None and e["crc32c"] != masked_crc32c(raw):
            raise IOError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, _DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def to_params(tensors: dict) -> "OrderedDict[str, np.ndarray]":
    """Checkpoint names -> graph tensor names (`conv2d/kernel` -> `conv2d/kernel:0`), the keys of the `params=` dicts."""
    return OrderedDict((k + ":0", v) for k, v in tensors.items())


# ------------------------------------------------------------------------------------------------- writer
def _block(entries) -> bytes:
    """One table block without prefix compression (every entry is a restart point: valid for any reader)."""
    body, restarts = b"", []
    for key, val in entries:
        restarts.append(len(body))
        body += _varint(0) + _varint(len(key)) + _varint(len(val)) + key + val
    if not restarts:
        restarts = [0]
    return body + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def _emit(buf: bytearray, block: bytes):
    off = len(buf)
    buf += block + b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00"))
    return off, len(block)


def save_checkpoint(prefix: str, tensors: dict) -> None:
    """Write `<prefix>.index` + `<prefix>.data-00000-of-00001` (names without ':0')."""
    names = sorted(tensors)
    data = bytearray()
    entries = [(b"", _varint((1 << 3) | 0) + _varint(1) + _varint((2 << 3) | 0) + _varint(0) +
                _ld(3, _varint((1 << 3) | 0) + _varint(1)))]  # num_shards=1, endianness=LITTLE, version{producer=1}
    for name in names:
        a = np.asarray(tensors[name])  # (ascontiguousarray would turn a scalar such as global_step into shape (1,))
        dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
        if np.dtype(dt) not in _DTYPE_IDS:
            raise NotImplementedError(f"{name}: dtype {a.dtype}")
        raw = a.astype(dt).tobytes(order="C")
        shape = b"".join(_ld(2, _varint((1 << 3) | 0) + _varint(d)) for d in a.shape)
        e = _varint((1 << 3) | 0) + _varint(_DTYPE_IDS[np.dtype(dt)]) + _ld(2, shape)
        e += _varint((4 << 3) | 0) + _varint(len(data)) + _varint((5 << 3) | 0) + _varint(len(raw))
        e += _varint((6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw))
        entries.append((name.encode(), e))
        data += raw
    buf = bytearray()
    doff, dsize = _emit(buf, _block(entries))
    moff, msize = _emit(buf, _block([]))
    ioff, isize = _emit(buf, _block([(entries[-1][0] + b"\x00", _varint(doff) + _varint(dsize))]))
    footer = _varint(moff) + _varint(msize) + _varint(ioff) + _varint(isize)
    buf += footer + bytes(40 - len(footer)) + struct.pack("<Q", MAGIC)
    open(prefix + ".index", "wb").write(bytes(buf))
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
