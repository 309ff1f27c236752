"""vdsr/vdsr/experiment_train.py of the reference on the B200 hot path: same flags, same loop
(step -> stepwise learning rate -> next(image_batches) -> session.run({'step','loss','trainer'}, feeds)), checkpoints as `.npz`
keyed by the TF variable names, the scalar summaries as JSON lines under --logs_path.
    python -m ml_super_resolution_b200.vdsr.experiment_train --data_path DIR --ckpt_path DIR --logs_path DIR [--config vdsr.yaml]"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from .. import flags
from ..io.images import imread_u8, list_images
from ..params import load_params
from ..session import Session, placeholder
from . import dataset, model_vdsr

FLAGS = flags.FLAGS


def build_dataset():
    """reference :37-54: an endless generator of (sd_images, hd_images) batches in [-1, 1]."""
    factors = [float(s) for s in FLAGS.scaling_factors.split("_")]
    images = [imread_u8(p) for p in list_images(FLAGS.data_path)]
    images = [im for im in images if min(im.shape[:2]) > FLAGS.image_size]
    if not images:
        raise SystemExit(f"no image larger than {FLAGS.image_size} px under {FLAGS.data_path}")
    return dataset.image_batches(images, factors, FLAGS.image_size, FLAGS.batch_size)


def build_model(params=None):
    """reference :11-34."""
    sd_images = placeholder([None, None, None, 3], "sd_images")
    hd_images = placeholder([None, None, None, 3], "hd_images")
    return model_vdsr.build_model(sd_images, hd_images, FLAGS.num_layers, FLAGS.use_adam, params=params)


def latest_checkpoint(ckpt_dir):
    """tf.train.latest_checkpoint for the `model.ckpt-<step>.npz` files this driver writes."""
    found = glob.glob(os.path.join(ckpt_dir, "model.ckpt-*.npz"))
    return max(found, key=lambda p: int(p.rsplit("-", 1)[1][:-4])) if found else None


def main(_):
    os.makedirs(FLAGS.ckpt_path, exist_ok=True)
    os.makedirs(FLAGS.logs_path, exist_ok=True)
    image_batches = build_dataset()
    source = latest_checkpoint(FLAGS.ckpt_path)
    model = build_model(load_params(source) if source else None)
    net = model["sr_images"].graph.net
    if source:
        net.step = int(np.load(source)["global_step"]) if "global_step" in np.load(source).files else 0
    log = open(os.path.join(FLAGS.logs_path, "events.jsonl"), "a")
    with Session() as session:
        while True:
            step = session.run(model["step"])
            if step == FLAGS.stop_training_at_k_step:
                net.arena.save(os.path.join(FLAGS.ckpt_path, f"model.ckpt-{step}.npz"), global_step=step)
                break
            lr = FLAGS.initial_learning_rate * (FLAGS.learning_rate_decay_factor ** (step // FLAGS.learning_rate_decay_steps))
            sd_images, hd_images = next(image_batches)
            feeds = {model["sd_images"]: sd_images, model["hd_images"]: hd_images, model["learning_rate"]: lr}
            fetch = {"step": model["step"], "loss": model["loss"], "trainer": model["trainer"]}
            if (step + 1) % 100 == 0:
                fetch["psnr"] = model["psnr"]  # the reference's 'epoch' summary (:78-91)
            fetched = session.run(fetch, feed_dict=feeds)
            rec = {"step": int(fetched["step"]), "loss": float(fetched["loss"]), "learning_rate": lr}
            if "psnr" in fetched:
                rec["psnr"] = float(np.mean(fetched["psnr"]))
            log.write(json.dumps(rec) + "\n")
    log.close()


if __name__ == "__main__":
    flags.DEFINE_string("data_path", None, "path to a directory which contains all image training data")
    flags.DEFINE_string("ckpt_path", None, "path to a directory for keeping the checkpoint")
    flags.DEFINE_string("logs_path", None, "path to a directory for keeping log")
    flags.DEFINE_string("scaling_factors", "2_3_4", "different scaling factors for training, separated by _")
    flags.DEFINE_integer("image_size", 41, "size of training images")
    flags.DEFINE_integer("batch_size", 64, "size of each batch during training")
    flags.DEFINE_integer("num_layers", 20, "number of hidden layers")
    flags.DEFINE_integer("learning_rate_decay_steps", 2560, "")
    flags.DEFINE_float("learning_rate_decay_factor", 0.1, "")
    flags.DEFINE_float("initial_learning_rate", 0.1, "")
    flags.DEFINE_integer("stop_training_at_k_step", 12800, "stop training at k step, default is stop after 80 epochs")
    flags.DEFINE_boolean("use_adam", True, "use adam instead of momentum optimizer")
    flags.run(main)
