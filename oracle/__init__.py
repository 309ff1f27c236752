"""CPU oracle for the convolutional hot path of imironhead/ml_super_resolution.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`ml_super_resolution_b200/`) may import this; the only legal importers are
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs.

PARITY UNPINNED, with one exception: the reference ships no tests, golden vectors
or fixtures and its arithmetic lives in un-vendored
TensorFlow 1.8 / scikit-image 0.14, which cannot be installed in this image.  The
oracle restates the published semantics of those ops and
is cross-validated internally (torch fp64 conv vs explicit numpy im2col GEMM vs
finite differences; bilinear vs cv2.INTER_LINEAR; SSIM vs scipy.ndimage filters;
pixel-shuffle vs the reference's own numpy pack/unpack code paths restated
verbatim; the tf.train.Example wire codec vs the protobuf runtime).
PINNED: Pillow IS present here, and `ops.pil_resize_u8` -- what
`scipy.misc.imresize` does in EnhanceNet's input pipeline -- is checked bit for
bit against it (tests/test_oracle_cpu.py); the gaussian of the degrade pre-pass is
checked to 1e-12 against `scipy.ndimage.gaussian_filter(mode='nearest',
truncate=4.0)`, the very call `skimage.filters.gaussian` delegates to.
"""
from . import ops, models  # noqa: F401
