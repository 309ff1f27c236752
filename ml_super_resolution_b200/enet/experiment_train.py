"""enet/enet/experiment_train.py of the reference: same flags and loop -- every step trains the generator, every third step the
discriminator first (:110-158), checkpoints every 1000 steps -- on the device-resident input pipeline (`datasets.image_batches`)
and the B200 losses (`build_enet`).  Scalar summaries go to JSON lines under --log_path, checkpoints are `.npz` files keyed by
the TF variable names (generator `g_/...`, discriminator `d_/...`).
    python -m ml_super_resolution_b200.enet.experiment_train --train_dir_path DIR --vgg19_path vgg19.npz --ckpt_path DIR --log_path DIR"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from .. import flags
from ..io.images import imread_u8, list_images
from ..params import load_params
from ..session import Session, placeholder
from . import datasets, model_enet

FLAGS = flags.FLAGS


def build_training_model(params=None, d_params=None):
    """reference :11-33."""
    sd_images = placeholder([None, 32, 32, 3], "sd_images")
    bq_images = placeholder([None, 128, 128, 3], "bq_images")
    hd_images = placeholder([None, 128, 128, 3], "hd_images")
    model = model_enet.build_enet(sd_images, bq_images, hd_images, FLAGS.model, FLAGS.vgg19_path, params=params, d_params=d_params)
    images = [im for im in (imread_u8(p) for p in list_images(FLAGS.train_dir_path)) if min(im.shape[:2]) >= 255]
    model["image_batches"] = datasets.image_batches(images, 4.0, batch_size=FLAGS.batch_size)
    return model


def save(trainer, step):
    arrays = dict(trainer.gen.arena.to_numpy())
    if trainer.disc is not None:
        arrays.update(trainer.disc.arena.to_numpy())
    np.savez(os.path.join(FLAGS.ckpt_path, f"model.ckpt-{step}.npz"), global_step=step, **arrays)


def main(_):
    if FLAGS.model not in ["p", "pa", "pat"]:
        FLAGS.model = "pat"
    os.makedirs(FLAGS.ckpt_path, exist_ok=True)
    os.makedirs(FLAGS.log_path, exist_ok=True)
    found = glob.glob(os.path.join(FLAGS.ckpt_path, "model.ckpt-*.npz"))
    source = max(found, key=lambda p: int(p.rsplit("-", 1)[1][:-4])) if found else None
    g_params = d_params = None
    if source:
        allp = load_params(source)
        g_params = {k: v for k, v in allp.items() if k.startswith("g_")}
        d_params = {k: v for k, v in allp.items() if k.startswith("d_")} or None
    model = build_training_model(g_params, d_params)
    trainer = model["sr_images"].graph.t
    if source:
        trainer.step = int(source.rsplit("-", 1)[1][:-4])
    log = open(os.path.join(FLAGS.log_path, "events.jsonl"), "a")
    stop = FLAGS.stop_training_at_k_step or None  # (extension: the reference loops until interrupted)
    with Session() as session:
        while True:
            step = session.run(model["step"])
            if step % 1000 == 999 or step == stop:
                save(trainer, step)
            if step == stop:
                break
            if step % 3 == 0 and "d_trainer" in model:  # train discriminator
                sd, bq, hd = next(model["image_batches"])
                f = session.run({"step": model["step"], "trainer": model["d_trainer"], "a_loss": model["a_loss"]},
                                feed_dict={model["sd_images"]: sd, model["bq_images"]: bq, model["hd_images"]: hd})
                log.write(json.dumps({"step": int(step), "discriminator_loss": f["a_loss"]}) + "\n")
            sd, bq, hd = next(model["image_batches"])  # train generator
            fetch = {"step": model["step"], "trainer": model["g_trainer"], "perceptual_loss": model["p_loss"]}
            if "g_loss" in model:
                fetch["generator_loss"] = model["g_loss"]
            if "t_loss" in model:
                fetch["texture_loss"] = model["t_loss"]
            f = session.run(fetch, feed_dict={model["sd_images"]: sd, model["bq_images"]: bq, model["hd_images"]: hd})
            log.write(json.dumps({"step": int(step), **{k: v for k, v in f.items() if k.endswith("_loss")}}) + "\n")
    log.close()


if __name__ == "__main__":
    flags.DEFINE_string("train_dir_path", None, "")
    flags.DEFINE_string("vgg19_path", None, "")
    flags.DEFINE_string("ckpt_path", None, "")
    flags.DEFINE_string("log_path", None, "")
    flags.DEFINE_string("model", "pat", "")
    flags.DEFINE_integer("batch_size", 64, "")
    flags.DEFINE_integer("stop_training_at_k_step", 0, "stop after k generator steps (0: run until interrupted, as the reference does)")
    flags.run(main)
