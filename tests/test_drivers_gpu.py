"""The experiment drivers end to end on tiny settings (subprocesses, like `python -m vdsr.experiment_train ...` in the
reference's makefiles): flags -> dataset -> session loop -> checkpoint -> evaluate / resolve."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(module, *args):
    r = subprocess.run([sys.executable, "-m", module, *args], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def _images(d, n=3, size=96, ext="png"):
    from PIL import Image
    rng = np.random.default_rng(1)
    os.makedirs(d, exist_ok=True)
    for i in range(n):
        base = rng.integers(0, 256, (size // 8, size // 8, 3), dtype=np.uint8)
        Image.fromarray(base).resize((size, size), Image.BICUBIC).save(os.path.join(d, f"im{i}.{ext}"))


def test_vdsr_train_evaluate_resolve(tmp_path):
    data, ckpt, logs = str(tmp_path / "data"), str(tmp_path / "ckpt"), str(tmp_path / "logs")
    _images(data)
    _run("ml_super_resolution_b200.vdsr.experiment_train", f"--data_path={data}", f"--ckpt_path={ckpt}", f"--logs_path={logs}",
         "--num_layers=4", "--batch_size=8", "--stop_training_at_k_step=3", "--initial_learning_rate=0.001")
    ck = os.path.join(ckpt, "model.ckpt-3.npz")
    assert os.path.exists(ck) and len(open(os.path.join(logs, "events.jsonl")).read().splitlines()) == 3
    out = _run("ml_super_resolution_b200.vdsr.experiment_evaluate", f"--ckpt_path={ck}", f"--hd_image_dir_path={data}", "--scaling_factor=3")
    assert "psnr (sd, sr):" in out and "ssim (sd, sr):" in out
    png = str(tmp_path / "sr.png")
    out = _run("ml_super_resolution_b200.vdsr.experiment_resolve", f"--ckpt_path={ck}", f"--hd_image_path={data}/im0.png", f"--sr_image_path={png}",
               "--scaling_factor=2")
    from PIL import Image
    assert Image.open(png).size == (96, 96) and "psnr(sd, sr):" in out


def test_espcn_test_driver(tmp_path):
    from oracle import models as OM
    data = str(tmp_path / "data")
    _images(data, n=2, size=60)
    p = OM.espcn_init(seed=1, scaling_factor=3, channels=3)
    ck = str(tmp_path / "espcn.npz")
    np.savez(ck, **p)
    out = _run("ml_super_resolution_b200.espcn.experiment_test", f"--data_path={data}", f"--ckpt_path={ck}", "--score_space=y")
    assert out.count("psnr:") == 3 and "ssim:" in out
    res = str(tmp_path / "hr.png")
    _run("ml_super_resolution_b200.espcn.experiment_test", f"--data_path={data}/im0.png", f"--ckpt_path={ck}", f"--result_path={res}")
    from PIL import Image
    assert Image.open(res).size == (180, 180)


def test_srcnn_main_trains_from_jpegs(tmp_path):
    data, ckpt, logs = str(tmp_path / "jpg"), str(tmp_path / "ckpt"), str(tmp_path / "logs")
    _images(data, n=2, size=80, ext="jpg")
    _run("ml_super_resolution_b200.srcnn.srcnn", "--train", f"--training-images-path={data}", f"--ckpt-dir-path={ckpt}", f"--logs-dir-path={logs}",
         "--batch-size=4", "--crop-image-size=66", "--stop_training_at_k_step=2")
    assert os.path.exists(os.path.join(ckpt, "model.ckpt-2.npz"))


def test_enet_train_and_resolve(tmp_path):
    data, ckpt, logs = str(tmp_path / "data"), str(tmp_path / "ckpt"), str(tmp_path / "logs")
    _images(data, n=2, size=256)
    _run("ml_super_resolution_b200.enet.experiment_train", f"--train_dir_path={data}", f"--ckpt_path={ckpt}", f"--log_path={logs}", "--model=pat",
         "--batch_size=2", "--stop_training_at_k_step=2")
    ck = os.path.join(ckpt, "model.ckpt-2.npz")
    assert os.path.exists(ck)
    keys = np.load(ck).files
    assert any(k.startswith("g_/") for k in keys) and any(k.startswith("d_/") for k in keys)
    small = str(tmp_path / "small")
    _images(small, n=1, size=40)
    gen = str(tmp_path / "gen.npz")
    _run("ml_super_resolution_b200.enet.experiment_resolve", "--extract_model", f"--source_ckpt_path={ck}", f"--target_ckpt_path={gen}")
    out = str(tmp_path / "out")
    _run("ml_super_resolution_b200.enet.experiment_resolve", f"--graph_define_path={gen}", f"--source_dir_path={small}", f"--target_dir_path={out}")
    from PIL import Image
    assert Image.open(os.path.join(out, "im0_sr.png")).size == (160, 160) and os.path.exists(os.path.join(out, "im0_bq.png"))
