"""Host logic of the experiment drivers that needs no GPU: the tf.app.flags stand-in (names, defaults, --config overlay,
boolean spellings) and the image hand-off helpers."""
import numpy as np
import pytest

from ml_super_resolution_b200 import flags


def _fresh():
    flags.FLAGS._defs.clear()
    flags.FLAGS._vals.clear()


def test_flags_defaults_and_command_line():
    _fresh()
    flags.DEFINE_string("data_path", None, "")
    flags.DEFINE_integer("batch_size", 64, "")
    flags.DEFINE_float("initial_learning_rate", 0.1, "")
    flags.DEFINE_boolean("use_adam", True, "")
    f = flags.parse(["--data_path=/x/y-z", "--batch_size", "32", "--nouse_adam"])
    assert f.data_path == "/x/y-z" and f.batch_size == 32 and f.initial_learning_rate == 0.1 and f.use_adam is False
    f = flags.parse(["--use_adam=true", "--initial_learning_rate=1e-3"])
    assert f.use_adam is True and f.initial_learning_rate == pytest.approx(1e-3)
    with pytest.raises(SystemExit):
        flags.parse(["--no_such_flag=1"])


def test_flags_config_yaml_overlay(tmp_path):
    _fresh()
    flags.DEFINE_integer("num_layers", 20, "")
    flags.DEFINE_string("scaling_factors", "2_3_4", "")
    cfg = tmp_path / "config.yaml"
    cfg.write_text("trainingInput:\n    scaleTier: CUSTOM\n    masterType: standard_p100\nnum_layers: 8\nscaling_factors: '3'\n")
    f = flags.parse(["--config", str(cfg), "--scaling_factors=2_3"])  # the command line wins over the file
    assert f.num_layers == 8 and f.scaling_factors == "2_3"


def test_flags_run_calls_main():
    _fresh()
    flags.DEFINE_integer("k", 1, "")
    seen = []
    flags.run(lambda _: seen.append(flags.FLAGS.k), ["--k=7"])
    assert seen == [7]


def test_image_helpers_roundtrip(tmp_path):
    from ml_super_resolution_b200.io.images import imread_u8, list_images, write_png
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    write_png(str(tmp_path / "b.png"), a)
    write_png(str(tmp_path / "a.png"), a[..., :1])
    (tmp_path / "notes.txt").write_text("x")
    paths = list_images(str(tmp_path))
    assert [p.rsplit("/", 1)[1] for p in paths] == ["a.png", "b.png"]
    assert np.array_equal(imread_u8(paths[1]), a)
    assert np.array_equal(imread_u8(paths[0]), np.repeat(a[..., :1], 3, axis=2))
