"""espcn/espcn/experiment_test.py of the reference: --data_path a directory -> PSNR / SSIM per image (scored in PACKED space,
RGB or Y, :30-53); --data_path a file -> the super-resolved image written to --result_path.  The un-pack the reference does on
the host (np.split / reshape / concatenate, :173-177) and its [0,1] clip + uint8 conversion are fused into the one kernel."""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import flags, metrics
from ..io.images import imread_u8, list_images, write_png
from ..session import Session
from . import model_espcn

FLAGS = flags.FLAGS


def build_model():
    """reference :14-55 (the psnr / ssim fetches are `evaluate_pair` below)."""
    return model_espcn.build_test_model(FLAGS.ckpt_path + ".meta", FLAGS.ckpt_path)


def gaussian_nearest(image, sigma):
    """skimage.filters.gaussian(image, sigma, mode='nearest') over the two spatial axes: scipy.ndimage.gaussian_filter with
    truncate = 4.0, the call skimage delegates to."""
    from scipy import ndimage
    return ndimage.gaussian_filter(image, [sigma, sigma, 0], mode="nearest", truncate=4.0)


def prepare_image_pair(hr_image_path, upscaling_factor):
    """reference :58-101: trim to a multiple of the factor, [-1,1], gaussian blur, decimate from offset factor//2; the target
    is returned in packed space."""
    hr_image = imread_u8(hr_image_path)
    h, w, _ = hr_image.shape
    h -= h % upscaling_factor
    w -= w % upscaling_factor
    hr_image = hr_image[:h, :w] / 127.5 - 1.0
    sigma = np.maximum(0.0, 0.5 * (upscaling_factor - 1.0))
    bl_image = gaussian_nearest(hr_image, sigma)
    offset = upscaling_factor // 2
    lr_image = bl_image[offset::upscaling_factor, offset::upscaling_factor]
    hr_patches = np.split(hr_image, w // upscaling_factor, axis=1)
    hr_patches = [np.reshape(im, [h // upscaling_factor, 1, -1]) for im in hr_patches]
    return lr_image.astype(np.float32), np.concatenate(hr_patches, axis=1).astype(np.float32)


def evaluate_images():
    model = build_model()
    psnrs, ssims = [], []
    with Session() as session:
        for image_path in list_images(FLAGS.data_path):
            lr_image, hr_image = prepare_image_pair(image_path, model["scaling_factor"])
            sr = session.run(model["sr_results"], feed_dict={model["lr_sources"]: np.expand_dims(lr_image, 0)})
            p, s = metrics.espcn_scores(torch.from_numpy(sr).cuda(), torch.from_numpy(hr_image[None]).cuda(), model["scaling_factor"], FLAGS.score_space)
            psnrs.append(float(p[0]))
            ssims.append(float(s[0]))
            print("name: {:>32}, psnr: {:.4f}, ssim: {:.4f}".format(os.path.basename(image_path), psnrs[-1], ssims[-1]))
    print("data: {}".format(FLAGS.data_path))
    print("psnr: {0:.4f}".format(np.mean(psnrs)))
    print("ssim: {0:.4f}".format(np.mean(ssims)))


def super_resolve_image():
    model = build_model()
    lr_image = (imread_u8(FLAGS.data_path) / 127.5 - 1.0).astype(np.float32)
    with Session() as session:
        # uint8 = saturate_cast(sr * 127.5 + 127.5) == the reference's clip(sr * 0.5 + 0.5, 0, 1) handed to skimage.io.imsave
        image = session.run(model["hr_images_u8"], feed_dict={model["lr_sources"]: np.expand_dims(lr_image, 0)})
    write_png(FLAGS.result_path, image[0])


def main(_):
    if os.path.isdir(FLAGS.data_path):
        evaluate_images()
    else:
        super_resolve_image()


if __name__ == "__main__":
    flags.DEFINE_string("data_path", None, "path to the test data directory")
    flags.DEFINE_string("ckpt_path", None, "path to the checkpoint")
    flags.DEFINE_string("result_path", None, "path for the super-resolved image")
    flags.DEFINE_string("score_space", "y", "evaluate on y(uv) or rgb")
    flags.run(main)
