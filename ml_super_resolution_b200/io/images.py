"""Image file hand-off of the experiment drivers (the reference goes through skimage.io / tf.gfile / tf.image.encode_png):
decoded uint8 HWC arrays in, PNG files out, through Pillow (the library scipy.misc / skimage delegate to as well)."""
from __future__ import annotations

import os

import numpy as np

EXTENSIONS = (".png", ".jpg", ".bmp", ".jpeg")


def list_images(dir_path: str):
    """vdsr/vdsr/experiment_evaluate.py:70-76: names with an image extension, joined to the directory."""
    return [os.path.join(dir_path, n) for n in sorted(os.listdir(dir_path)) if n.lower().endswith(EXTENSIONS)]


def imread_u8(path: str) -> np.ndarray:
    """uint8 [H,W,3] (grey images are replicated, alpha is dropped), what skimage.io.imread + the drivers' shape checks keep."""
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


def write_png(path: str, image_u8: np.ndarray) -> None:
    """tf.image.encode_png of a uint8 [H,W,C] image (vdsr/vdsr/experiment_resolve.py:65-69,124-125)."""
    from PIL import Image
    a = np.asarray(image_u8, dtype=np.uint8)
    Image.fromarray(a[..., 0] if a.ndim == 3 and a.shape[2] == 1 else a).save(path, format="PNG")
