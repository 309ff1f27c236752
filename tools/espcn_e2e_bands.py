"""Development timing of EspcnNet.forward_host (the Session.run host path) for different row-band sizes, raw uint8 frames in /
uint8 frames out, 4 x 1080p-LR Y frames per call."""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200.espcn.model_espcn import EspcnNet  # noqa: E402
from ml_super_resolution_b200.session import pinned_empty  # noqa: E402

net = EspcnNet(None, 3, 1)
raw = pinned_empty((4, 1080, 1920, 1), "uint8")
raw[...] = np.random.default_rng(0).integers(0, 256, raw.shape, dtype=np.uint8)
out = pinned_empty((4, 3240, 5760, 1), "uint8")
rt, ot = torch.from_numpy(raw), torch.from_numpy(out)
import time
for band in (135, 270, 540, 1080):
    for _ in range(3):
        net.forward_host(rt, ot, True, True, band_rows=band)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        net.forward_host(rt, ot, True, True, band_rows=band)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"band_rows {band:5d}: {ms:.3f} ms per call  {4 * 3240 * 5760 / ms / 1e6:.1f} Gpix/s", flush=True)
