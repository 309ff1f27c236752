"""vdsr/vdsr/experiment_evaluate.py of the reference: PSNR / SSIM of (hd, sd) and (hd, sr) over a directory of images.
--meta_path is accepted and ignored (the graph is rebuilt from the variable shapes); --ckpt_path is an `.npz` keyed by the TF
variable names or a TensorFlow Saver checkpoint prefix."""
from __future__ import annotations

import time

import numpy as np
import torch

from .. import flags, metrics
from ..io.images import imread_u8, list_images
from ..params import load_params
from ..session import Session, placeholder
from . import dataset, model_vdsr

FLAGS = flags.FLAGS


def load_image(hd_image_path, scaling_factor):
    """reference :14-35."""
    hd_image = imread_u8(hd_image_path).astype(np.float32) / np.float32(255.0)  # skimage.util.img_as_float32
    sd_image = dataset.hd_image_to_sd_image(hd_image, scaling_factor)
    return np.expand_dims(sd_image * 2.0 - 1.0, 0), np.expand_dims(hd_image * 2.0 - 1.0, 0)


def build_model():
    """reference :38-61: restored weights + the four tf.image.psnr / ssim fetches (max_val 2.0)."""
    params = load_params(FLAGS.ckpt_path)
    num_layers = sum(1 for k in params if k.endswith("kernel:0"))
    sd_ph, hd_ph = placeholder([None, None, None, 3], "sd_images"), placeholder([None, None, None, 3], "hd_images")
    model = model_vdsr.build_model(sd_ph, hd_ph, num_layers, params=params)
    return model


def scores(sr, sd, hd):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()  # noqa: E731
    sr, sd, hd = t(sr), t(sd), t(hd)
    return {"hd_sd_psnrs": metrics.psnr(hd, sd, 2.0).cpu().numpy(), "hd_sr_psnrs": metrics.psnr(hd, sr, 2.0).cpu().numpy(),
            "hd_sd_ssims": metrics.ssim(hd, sd, 2.0).cpu().numpy(), "hd_sr_ssims": metrics.ssim(hd, sr, 2.0).cpu().numpy()}


def main(_):
    image_paths = list_images(FLAGS.hd_image_dir_path)
    acc = {k: [] for k in ("hd_sd_psnrs", "hd_sr_psnrs", "hd_sd_ssims", "hd_sr_ssims")}
    total_time = 0.0
    model = build_model()
    with Session() as session:
        for image_path in image_paths:  # one by one: the image sizes differ
            sd_images, hd_images = load_image(image_path, FLAGS.scaling_factor)
            begin_time = time.time()
            sr_images = session.run(model["sr_images"], feed_dict={model["sd_images"]: sd_images})
            fetched = scores(sr_images, sd_images, hd_images)
            total_time += time.time() - begin_time
            for k in acc:
                acc[k].append(fetched[k][0])
    print("x{}".format(FLAGS.scaling_factor))
    print("time (s)     : {}".format(total_time / max(1, len(image_paths))))
    print("psnr (sd, sr): {}, {}".format(np.mean(acc["hd_sd_psnrs"]), np.mean(acc["hd_sr_psnrs"])))
    print("ssim (sd, sr): {}, {}".format(np.mean(acc["hd_sd_ssims"]), np.mean(acc["hd_sr_ssims"])))


if __name__ == "__main__":
    flags.DEFINE_string("meta_path", None, "path to the graph (accepted, unused)")
    flags.DEFINE_string("ckpt_path", None, "path to the weights")
    flags.DEFINE_string("hd_image_dir_path", None, "path to a directory of ground-truth images")
    flags.DEFINE_integer("scaling_factor", 2, "scaling factor of the super-resolution task")
    flags.run(main)
