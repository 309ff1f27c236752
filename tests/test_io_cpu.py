"""the reference's ESPCN patch-pair TFRecords, read and written without TensorFlow."""
import os
import struct

import numpy as np
import pytest

from ml_super_resolution_b200.io import tfrecord as T


def test_crc32c_known_answers():
    # CRC-32C check value (RFC 3720 appendix B.4) and the all-zero / all-one 32-byte vectors from the same appendix
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    c = T.crc32c(b"abc")
    assert T.masked_crc32c(b"abc") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_wire_encoding_is_the_published_protobuf_layout():
    # hand-assembled tf.train.Example {features {feature {key: "a" value {int64_list {value: 5}}}}}
    feat = bytes([0x1A, 0x03, 0x0A, 0x01, 0x05])              # Feature.int64_list(3) -> Int64List.value(1, packed) = [5]
    entry = bytes([0x0A, 0x01]) + b"a" + bytes([0x12, len(feat)]) + feat
    features = bytes([0x0A, len(entry)]) + entry
    example = bytes([0x0A, len(features)]) + features
    assert T.encode_example({"a": 5}) == example
    assert T.decode_example(example) == {"a": 5}
    # unpacked int64 (older writers) and negative values decode too
    unpacked = bytes([0x0A, 0x0B, 0x0A, 0x09, 0x0A, 0x01]) + b"a" + bytes([0x12, 0x04, 0x1A, 0x02, 0x08, 0x07])
    assert T.decode_example(unpacked) == {"a": 7}
    assert T.decode_example(T.encode_example({"n": -3, "b": b"\x00\x01"})) == {"n": -3, "b": b"\x00\x01"}


def test_patch_pair_roundtrip_and_framing(tmp_path):
    rng = np.random.default_rng(0)
    r = 3
    paths = []
    pairs = []
    for i in range(5):
        lr = rng.standard_normal((17, 17, 3)).astype(np.float32)
        hr = rng.standard_normal((17, 17, 3 * r * r)).astype(np.float32)
        p = os.path.join(tmp_path, f"p{i}.tfrecord")
        T.write_patch(p, lr, hr)
        paths.append(p)
        pairs.append((lr, hr))
    raw = open(paths[0], "rb").read()
    (n,) = struct.unpack("<Q", raw[:8])
    assert len(raw) == 8 + 4 + n + 4 and struct.unpack("<I", raw[8:12])[0] == T.masked_crc32c(raw[:8])
    recs = list(T.read_records(paths[0]))
    assert len(recs) == 1
    lr, hr = T.decode_patch_pair(recs[0], r)
    assert np.array_equal(lr, pairs[0][0]) and np.array_equal(hr, pairs[0][1])
    # corruption is detected
    bad = bytearray(raw)
    bad[20] ^= 0xFF
    open(paths[1], "wb").write(bytes(bad))
    with pytest.raises(IOError):
        list(T.read_records(paths[1]))
    open(paths[1], "wb").write(raw)
    # batches: every record appears once per epoch
    it = T.patch_batches(str(tmp_path), batch_size=5, upscaling_factor=r, seed=1)
    lrb, hrb = next(it)
    assert lrb.shape == (5, 17, 17, 3) and hrb.shape == (5, 17, 17, 27) and lrb.dtype == np.float32
    sums = sorted(float(x.sum()) for x in lrb)
    expect = sorted([float(pairs[0][0].sum())] * 2 + [float(p[0].sum()) for p in pairs[2:]])
    assert np.allclose(sums, expect)


def test_example_encoding_matches_the_protobuf_runtime():
    """Pin the hand-written wire codec against Google's protobuf runtime on the published tf.train.Example schema
    (tensorflow/core/example/{example,feature}.proto), built here from a descriptor: no TensorFlow needed."""
    pb = pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="example_test.proto", package="tft", syntax="proto3")

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, number, ftype, label=1, type_name=None, packed=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = type_name
        if packed is not None:
            f.options.packed = packed
        return f

    F = descriptor_pb2.FieldDescriptorProto
    field(msg("BytesList"), "value", 1, F.TYPE_BYTES, 3)
    field(msg("FloatList"), "value", 1, F.TYPE_FLOAT, 3, packed=True)
    field(msg("Int64List"), "value", 1, F.TYPE_INT64, 3, packed=True)
    feat = msg("Feature")
    feat.oneof_decl.add().name = "kind"
    for n, num, t in (("bytes_list", 1, ".tft.BytesList"), ("float_list", 2, ".tft.FloatList"), ("int64_list", 3, ".tft.Int64List")):
        field(feat, n, num, F.TYPE_MESSAGE, 1, t).oneof_index = 0
    feats = msg("Features")
    entry = feats.nested_type.add()
    entry.name = "FeatureEntry"
    entry.options.map_entry = True
    field(entry, "key", 1, F.TYPE_STRING)
    field(entry, "value", 2, F.TYPE_MESSAGE, 1, ".tft.Feature")
    field(feats, "feature", 1, F.TYPE_MESSAGE, 3, ".tft.Features.FeatureEntry")
    field(msg("Example"), "features", 1, F.TYPE_MESSAGE, 1, ".tft.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    Example = message_factory.GetMessageClass(pool.FindMessageTypeByName("tft.Example"))
    rng = np.random.default_rng(3)
    lr = rng.standard_normal((4, 5, 3)).astype(np.float32)
    ex = Example()
    ex.features.feature["lr_pixels"].bytes_list.value.append(lr.tobytes())
    ex.features.feature["lr_height"].int64_list.value.append(4)
    ex.features.feature["lr_width"].int64_list.value.append(5)
    ex.features.feature["neg"].int64_list.value.append(-7)
    wire = ex.SerializeToString(deterministic=True)
    ours = T.encode_example({"lr_pixels": lr.tobytes(), "lr_height": 4, "lr_width": 5, "neg": -7})
    assert ours == wire
    back = T.decode_example(wire)
    assert back["lr_height"] == 4 and back["lr_width"] == 5 and back["neg"] == -7 and back["lr_pixels"] == lr.tobytes()
    parsed = Example.FromString(ours)
    assert parsed.features.feature["lr_width"].int64_list.value[0] == 5


def test_tf_checkpoint_bundle_roundtrip_and_layout(tmp_path):
    """TensorFlow V2 checkpoint (tensor bundle) reader / writer: LevelDB-style table with footer magic, prefix-compressed
    keys, per-block and per-tensor masked crc32c.  Parity is unpinned (no TF-written file here): the reader is exercised on
    this module's writer and on a hand-assembled prefix-compressed block."""
    from ml_super_resolution_b200.io import tf_checkpoint as K
    rng = np.random.default_rng(0)
    tensors = {"conv2d/kernel": rng.standard_normal((3, 3, 3, 64)).astype(np.float32), "conv2d/bias": np.zeros(64, np.float32),
               "conv2d_1/kernel": rng.standard_normal((3, 3, 64, 64)).astype(np.float32), "global_step": np.asarray(1234, np.int64),
               "beta1_power": np.asarray(0.9, np.float32)}
    prefix = str(tmp_path / "model.ckpt-1234")
    K.save_checkpoint(prefix, tensors)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 48
    back = K.load_checkpoint(prefix)
    assert list(back) == sorted(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v)
    assert list(K.to_params(back))[0].endswith(":0")
    # a corrupted tensor byte is caught by the per-tensor checksum
    d = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    d[10] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(d))
    with pytest.raises(IOError):
        K.load_checkpoint(prefix)
    # prefix-compressed entries with a single restart point (what TableBuilder emits inside a restart interval)
    def entry(shared, suffix, val):
        return bytes([shared, len(suffix), len(val)]) + suffix + val
    block = entry(0, b"f1/bias", b"A") + entry(3, b"kernel", b"BB") + entry(1, b"2/bias", b"C")
    block += struct.pack("<I", 0) + struct.pack("<I", 1)
    assert list(K._block_entries(block)) == [(b"f1/bias", b"A"), (b"f1/kernel", b"BB"), (b"f2/bias", b"C")]


def test_load_params_accepts_npz_and_tf_checkpoint(tmp_path):
    """`extract_weights` / `build_test_model` take either an .npz keyed by TF tensor names or a Saver checkpoint prefix."""
    from ml_super_resolution_b200.io import tf_checkpoint as K
    from ml_super_resolution_b200.params import load_params
    rng = np.random.default_rng(2)
    w = {"f1/kernel:0": rng.standard_normal((5, 5, 3, 64)).astype(np.float32), "f1/bias:0": rng.standard_normal(64).astype(np.float32)}
    np.savez(str(tmp_path / "w.npz"), **w)
    K.save_checkpoint(str(tmp_path / "model.ckpt-7"), {k[:-2]: v for k, v in w.items()})
    for path in (str(tmp_path / "w.npz"), str(tmp_path / "model.ckpt-7")):
        got = load_params(path)
        assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    with pytest.raises(FileNotFoundError):
        load_params(str(tmp_path / "nothing"))
