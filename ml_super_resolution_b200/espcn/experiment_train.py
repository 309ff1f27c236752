"""espcn/espcn/experiment_train.py of the reference: same flags and loop (stepwise learning rate fed every step, :99-107),
training data = the reference's `*.tfrecord` patch pairs read without TensorFlow, checkpoints as `.npz` keyed by the TF
variable names, the loss summary as JSON lines under --logs_path.
    python -m ml_super_resolution_b200.espcn.experiment_train --data_path DIR --ckpt_path DIR --logs_path DIR"""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from .. import flags
from ..params import load_params
from ..session import Session, placeholder
from . import dataset, model_espcn

FLAGS = flags.FLAGS


def build_dataset_iterator():
    """reference :11-18."""
    return dataset.build_image_batch_iterator(FLAGS.data_path, FLAGS.batch_size, FLAGS.scaling_factor)


def build_model(params=None):
    """reference :21-29 (placeholders instead of the tf.data iterator's tensors)."""
    lr_source = placeholder([None, FLAGS.lr_patch_size, FLAGS.lr_patch_size, 3], "lr_source")
    hr_target = placeholder([None, FLAGS.lr_patch_size, FLAGS.lr_patch_size, 3 * FLAGS.scaling_factor ** 2], "hr_target")
    return model_espcn.build_model(lr_source, FLAGS.scaling_factor, hr_target, params=params)


def main(_):
    os.makedirs(FLAGS.ckpt_path, exist_ok=True)
    os.makedirs(FLAGS.logs_path, exist_ok=True)
    batches = build_dataset_iterator()
    found = glob.glob(os.path.join(FLAGS.ckpt_path, "model.ckpt-*.npz"))
    source = max(found, key=lambda p: int(p.rsplit("-", 1)[1][:-4])) if found else None
    model = build_model(load_params(source) if source else None)
    net = model["sr_result"].graph.net
    net.step = int(np.load(source)["global_step"]) if source and "global_step" in np.load(source).files else 0
    log = open(os.path.join(FLAGS.logs_path, "events.jsonl"), "a")
    with Session() as session:
        step = net.step
        while step < FLAGS.stop_training_at_k_step:
            lr_level = step // FLAGS.learning_rate_decay_steps
            lr_batch, hr_batch = next(batches)
            feeds = {model["lr_source"]: lr_batch, model["hr_target"]: hr_batch,
                     model["learning_rate"]: FLAGS.initial_learning_rate * (FLAGS.learning_rate_decay_factor ** lr_level)}
            fetched = session.run({"step": model["step"], "optimizer": model["optimizer"], "loss": model["loss"]}, feed_dict=feeds)
            step = fetched["step"] + 1
            log.write(json.dumps({"step": step, "loss": float(fetched["loss"])}) + "\n")
    log.close()
    net.arena.save(os.path.join(FLAGS.ckpt_path, f"model.ckpt-{step}.npz"), global_step=step)


if __name__ == "__main__":
    flags.DEFINE_string("data_path", None, "path to training data (tfrecord) directory")
    flags.DEFINE_string("ckpt_path", None, "path to a directory for keeping the checkpoint")
    flags.DEFINE_string("logs_path", None, "path to a directory for keeping log")
    flags.DEFINE_integer("batch_size", 64, "size of each batch during training")
    flags.DEFINE_integer("scaling_factor", 3, "scaling factor for training")
    flags.DEFINE_integer("lr_patch_size", 17, "size of lr_patch as training data")
    flags.DEFINE_float("initial_learning_rate", 0.1, "")
    flags.DEFINE_float("learning_rate_decay_factor", 0.1, "")
    flags.DEFINE_integer("learning_rate_decay_steps", 2560, "")
    flags.DEFINE_integer("stop_training_at_k_step", 10000, "stop training at k step")
    flags.run(main)
