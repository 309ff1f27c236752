"""enet/enet/experiment_resolve.py of the reference: super-resolve every image of --source_dir_path 4x into
<name>_bq.png (bicubic, Pillow = scipy.misc.imresize(.., 400, 'bicubic')) and <name>_sr.png (generator output), both
saturate_cast(x * 127.5 + 127.5).  --extract_model copies the generator's variables of --source_ckpt_path into
--target_ckpt_path (the reference freezes the graph there to save memory, :14-58; here the generator needs no graph file, so
--graph_define_path names that extracted `.npz`)."""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import flags, metrics
from ..io.images import imread_u8, write_png
from ..params import load_params
from ..session import Session, placeholder
from . import model_enet

FLAGS = flags.FLAGS


def extract_model():
    params = {k: v for k, v in load_params(FLAGS.source_ckpt_path).items() if k.startswith("g_")}
    np.savez(FLAGS.target_ckpt_path, **params)


def source_images():
    """reference :61-95."""
    from PIL import Image
    for fname in sorted(os.listdir(FLAGS.source_dir_path)):
        name, ext = os.path.splitext(fname)
        if ext.lower() not in [".png", ".jpg", ".jpeg"]:
            continue
        sd_u8 = imread_u8(os.path.join(FLAGS.source_dir_path, fname))
        h, w = sd_u8.shape[:2]
        bq_u8 = np.asarray(Image.fromarray(sd_u8).resize((4 * w, 4 * h), Image.BICUBIC))  # scipy.misc.imresize(sd, 400, 'bicubic')
        yield {"sd_image": np.expand_dims(sd_u8.astype(np.float32) / 127.5 - 1.0, 0), "bq_image": np.expand_dims(bq_u8.astype(np.float32) / 127.5 - 1.0, 0),
               "bq_path": os.path.join(FLAGS.target_dir_path, name + "_bq.png"), "sr_path": os.path.join(FLAGS.target_dir_path, name + "_sr.png")}


def super_resolve():
    params = {k: v for k, v in load_params(FLAGS.graph_define_path).items() if k.startswith("g_")}
    sd_ph, bq_ph = placeholder([None, None, None, 3], "sd_images"), placeholder([None, None, None, 3], "bq_images")
    sr = model_enet.build_generator(sd_ph, bq_ph, None, "g_", params=params)
    os.makedirs(FLAGS.target_dir_path, exist_ok=True)
    with Session() as session:
        for images in source_images():
            sr_image = session.run(sr, feed_dict={sd_ph: images["sd_image"], bq_ph: images["bq_image"]})
            for arr, path in ((sr_image[0], images["sr_path"]), (images["bq_image"][0], images["bq_path"])):
                write_png(path, metrics.saturate_cast_u8(torch.from_numpy(np.ascontiguousarray(arr, np.float32)).cuda()).cpu().numpy())


def main(_):
    if FLAGS.extract_model:
        extract_model()
    else:
        super_resolve()


if __name__ == "__main__":
    flags.DEFINE_boolean("extract_model", False, "")
    flags.DEFINE_string("source_ckpt_path", None, "")
    flags.DEFINE_string("target_ckpt_path", None, "")
    flags.DEFINE_string("graph_define_path", None, "")
    flags.DEFINE_string("source_dir_path", None, "")
    flags.DEFINE_string("target_dir_path", None, "")
    flags.run(main)
