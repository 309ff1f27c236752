"""vdsr/vdsr/experiment_resolve.py of the reference: super-resolve one image to a PNG
(saturate_cast(sr * 127.5 + 127.5) on the device, :65-69), optionally scoring it against its ground truth."""
from __future__ import annotations

import numpy as np
import torch

from .. import flags, metrics
from ..io.images import imread_u8, write_png
from ..session import Session
from . import dataset
from .experiment_evaluate import build_model, scores

FLAGS = flags.FLAGS


def load_image(hd_image_path, scaling_factor, ground_truth_mode):
    """reference :14-43: with ground truth the input is degraded from it; without, the image itself is the (already
    enlarged) low-resolution input."""
    hd_image = imread_u8(hd_image_path).astype(np.float32) / np.float32(255.0)
    sd_image = dataset.hd_image_to_sd_image(hd_image, scaling_factor) if ground_truth_mode else hd_image
    return np.expand_dims(sd_image * 2.0 - 1.0, 0), np.expand_dims(hd_image * 2.0 - 1.0, 0)


def main(_):
    sd_images, hd_images = load_image(FLAGS.hd_image_path, FLAGS.scaling_factor, FLAGS.ground_truth_mode)
    model = build_model()
    with Session() as session:
        sr_images = session.run(model["sr_images"], feed_dict={model["sd_images"]: sd_images})
    png = metrics.saturate_cast_u8(torch.from_numpy(sr_images[0]).cuda()).cpu().numpy()
    write_png(FLAGS.sr_image_path, png)
    if FLAGS.ground_truth_mode:
        f = scores(sr_images, sd_images, hd_images)
        print("psnr(sd, sr): {}, {}".format(f["hd_sd_psnrs"][0], f["hd_sr_psnrs"][0]))
        print("ssim(sd, sr): {}, {}".format(f["hd_sd_ssims"][0], f["hd_sr_ssims"][0]))


if __name__ == "__main__":
    flags.DEFINE_string("meta_path", None, "path to the graph (accepted, unused)")
    flags.DEFINE_string("ckpt_path", None, "path to the weights")
    flags.DEFINE_boolean("ground_truth_mode", True, "the input is a ground-truth image: degrade it first, report psnr / ssim")
    flags.DEFINE_string("hd_image_path", None, "path to the source image")
    flags.DEFINE_string("sr_image_path", None, "path to the result png")
    flags.DEFINE_float("scaling_factor", 2.0, "scaling factor of the super-resolution task")
    flags.run(main)
