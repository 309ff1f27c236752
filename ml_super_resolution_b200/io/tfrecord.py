"""ESPCN patch-pair TFRecords without TensorFlow -- the on-disk format written by `write_patch` and read by
`decode_patch_pair` / `build_image_batch_iterator` of espcn/espcn/dataset.py:10-49,160-195.

A `.tfrecord` file is a sequence of frames
    uint64 length | uint32 masked_crc32c(length) | bytes data[length] | uint32 masked_crc32c(data)     (little endian)
with masked_crc(x) = ((crc32c(x) >> 15 | crc32c(x) << 17) + 0xa282ead8) mod 2^32 (TensorFlow's record writer).  Each frame
holds one serialized `tf.train.Example`:
    Example   { Features features = 1; }
    Features  { map<string, Feature> feature = 1; }          // map entry: key = 1 (string), value = 2 (Feature)
    Feature   { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
    BytesList { repeated bytes value = 1; }   Int64List { repeated int64 value = 1 [packed]; }
The reference stores per record: lr_pixels / hr_pixels = raw little-endian fp32 bytes, lr|hr_{height,width,depth} = int64.
Only the protobuf wire encoding needed for that schema is implemented (varint, length-delimited); pure Python + numpy.
"""
from __future__ import annotations

import glob
import os
import struct

import numpy as np

# ------------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)
_TABLE = np.asarray(_TABLE, np.uint32)


def crc32c(data: bytes) -> int:
    crc = 0xFFFFFFFF
    tab = _TABLE
    for b in data:
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------- protobuf wire helpers
def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _ld(field: int, payload: bytes) -> bytes:  # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _fields(buf: bytes):
    """Yield (field number, wire type, value) of one message; value is int (varint) or bytes (length-delimited)."""
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


# ------------------------------------------------------------------------------------------------- tf.train.Example
def encode_example(features: dict) -> bytes:
    """features: name -> bytes (BytesList of one value) or int (Int64List of one value)."""
    entries = b""
    for name in sorted(features):  # protobuf map order is unspecified; TF's serializer emits sorted keys
        v = features[name]
        if isinstance(v, (bytes, bytearray)):
            feat = _ld(1, _ld(1, bytes(v)))
        else:
            feat = _ld(3, _ld(1, _varint(int(v))))  # packed repeated int64
        entries += _ld(1, _ld(1, name.encode()) + _ld(2, feat))
    return _ld(1, entries)


def decode_example(buf: bytes) -> dict:
    out = {}
    for f, _, features in _fields(buf):
        if f != 1:
            continue
        for g, _, entry in _fields(features):
            if g != 1:
                continue
            name, feat = None, None
            for k, _, v in _fields(entry):
                if k == 1:
                    name = v.decode()
                elif k == 2:
                    feat = v
            for kind, _, lst in _fields(feat):
                if kind == 1:  # BytesList
                    out[name] = [v for (i, _, v) in _fields(lst) if i == 1][0]
                elif kind == 3:  # Int64List, packed or not
                    vals = []
                    for i, wt, v in _fields(lst):
                        if i != 1:
                            continue
                        if wt == 0:
                            vals.append(v)
                        else:
                            p = 0
                            while p < len(v):
                                x, p = _read_varint(v, p)
                                vals.append(x)
                    x = vals[0]
                    out[name] = x - (1 << 64) if x >= (1 << 63) else x
                elif kind == 2:  # FloatList (not used by the reference's patches)
                    out[name] = np.frombuffer(b"".join(v for (i, _, v) in _fields(lst) if i == 1), "<f4")
    return out


# ------------------------------------------------------------------------------------------------- TFRecord framing
def write_records(path: str, payloads) -> None:
    with open(path, "wb") as f:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + data + struct.pack("<I", masked_crc32c(data)))


def read_records(path: str, verify: bool = True):
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if not head:
                return
            if len(head) != 8:
                raise IOError(f"{path}: truncated record header")
            (n,) = struct.unpack("<Q", head)
            (hc,) = struct.unpack("<I", f.read(4))
            data = f.read(n)
            tail = f.read(4)
            if len(data) != n or len(tail) != 4:
                raise IOError(f"{path}: truncated record")
            if verify and (hc != masked_crc32c(head) or struct.unpack("<I", tail)[0] != masked_crc32c(data)):
                raise IOError(f"{path}: record checksum mismatch")
            yield data


# ------------------------------------------------------------------------------------------------- the reference's patch pairs
def write_patch(patch_path: str, lr_patch: np.ndarray, hr_patch: np.ndarray) -> None:
    """espcn/espcn/dataset.py:177-195: one record per file, raw fp32 pixels + int64 shapes."""
    feat = {"lr_pixels": lr_patch.astype("<f4").tobytes(), "lr_height": lr_patch.shape[0], "lr_width": lr_patch.shape[1], "lr_depth": lr_patch.shape[2],
            "hr_pixels": hr_patch.astype("<f4").tobytes(), "hr_height": hr_patch.shape[0], "hr_width": hr_patch.shape[1], "hr_depth": hr_patch.shape[2]}
    write_records(patch_path, [encode_example(feat)])


def decode_patch_pair(record: bytes, scaling_factor: int = 3):
    """espcn/espcn/dataset.py:10-49: -> (lr_patch [h,w,3], hr_patch [H,W,3*r^2]) fp32.  Like the reference, the channel counts
    are fixed by the model (3 and 3*r^2), the stored depths are not used for the reshape."""
    f = decode_example(record)
    lr = np.frombuffer(f["lr_pixels"], "<f4").reshape(f["lr_height"], f["lr_width"], 3)
    hr = np.frombuffer(f["hr_pixels"], "<f4").reshape(f["hr_height"], f["hr_width"], 3 * scaling_factor ** 2)
    return lr, hr


def patch_batches(dir_path: str, batch_size: int = 32, upscaling_factor: int = 3, seed=None, shuffle_buffer: int = 100000):
    """espcn/espcn/dataset.py:52-92 `build_image_batch_iterator`: every `*.tfrecord` under `dir_path`, shuffled, repeated
    forever, batched.  Yields (lr_batch [B,h,w,3], hr_batch [B,H,W,3*r^2]) fp32 numpy arrays ready for `EspcnNet.train_step`
    (the shuffle uses a numpy generator: TensorFlow's shuffle order is not reproducible outside TF)."""
    paths = sorted(glob.glob(os.path.join(dir_path, "*.tfrecord")))
    if not paths:
        raise FileNotFoundError(f"no *.tfrecord under {dir_path}")
    rng = np.random.default_rng(seed)
    lrs, hrs = [], []
    while True:
        order = rng.permutation(len(paths))
        for i in order:
            for rec in read_records(paths[i]):
                lr, hr = decode_patch_pair(rec, upscaling_factor)
                lrs.append(lr)
                hrs.append(hr)
                if len(lrs) == batch_size:
                    yield np.stack(lrs), np.stack(hrs)
                    lrs, hrs = [], []
